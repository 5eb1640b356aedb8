// convertBinary_exe — prints a .bibd.gz file as text (ref: ASMC_SRC/SRC/main_convertBinary.cpp:6-23).
#include <iostream>

#include "BinaryDataReader.hpp"

int main(int argc, char* argv[])
{
  if (argc != 2) {
    std::cout << "Number of parameters is wrong." << std::endl;
    std::cout << "Only one parameter (name of binary file) is required." << std::endl;
    return 1;
  }
  BinaryDataReader reader(argv[1]);
  while (reader.moreLinesInFile()) {
    std::cout << reader.getNextLine().toString() << '\n';
  }
  return 0;
}
