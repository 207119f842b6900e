// fast variant: FMA contraction allowed, hoisted reciprocals, dry-step shortcuts.
#define MHM_FAST 1
#define MHM_KERNEL_NAME cell_block_kernel_fast
#define MHM_LAUNCH_NAME launch_cell_block_fast
#include "cell_kernel_launch.inc"
