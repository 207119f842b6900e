// strict variant: compiled with -fmad=false (see Makefile); literal formulas, IEEE division.
#define MHM_FAST 0
#define MHM_KERNEL_NAME cell_block_kernel_strict
#define MHM_LAUNCH_NAME launch_cell_block_strict
#include "cell_kernel_launch.inc"
