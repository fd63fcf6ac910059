/* procell_main.cpp - the `procell` executable: src/main.cu:18-34 of the reference, as one call into the C ABI. */
#include "../../include/procell_b200.h"

int main(int argc, char** argv) { return procell_main(argc, argv); }
