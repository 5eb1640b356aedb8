"""Drop-in for the reference's `asmc` Python package (ref: ASMC_SRC/SRC/__init__.py:18-33): the same class names,
served by the B200-native host library (fastsmc_b200/lib/pyASMC*.so over libfastsmc_b200.so)."""
import importlib.util
import glob
import os

_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


def _load():
    hits = sorted(glob.glob(os.path.join(_LIB, "pyASMC*.so")))
    if not hits:
        raise ImportError(f"pyASMC extension not built in {_LIB} — run `python -m fastsmc_b200.build` "
                          "(needs nvcc and g++); there is no CPU fallback")
    spec = importlib.util.spec_from_file_location("pyASMC", hits[-1])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


pyASMC = _load()

DecodingModeOverall = pyASMC.DecodingModeOverall
DecodingMode = pyASMC.DecodingMode
DecodingReturnValues = pyASMC.DecodingReturnValues
DecodePairsReturnStruct = pyASMC.DecodePairsReturnStruct
Individual = pyASMC.Individual
PairObservations = pyASMC.PairObservations
DecodingQuantities = pyASMC.DecodingQuantities
DecodingParams = pyASMC.DecodingParams
Data = pyASMC.Data
HMM = pyASMC.HMM
FastSMC = pyASMC.FastSMC
ASMC = pyASMC.ASMC
BinaryDataReader = pyASMC.BinaryDataReader
IbdPairDataLine = pyASMC.IbdPairDataLine
