// Test scaffolding: the reference includes this header but uses nothing from it (ref: ASMC_SRC/SRC/Data.cpp:32).
#pragma once
