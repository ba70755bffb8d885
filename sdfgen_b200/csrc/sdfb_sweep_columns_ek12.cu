// sdfb_sweep_columns_ek12.cu -- the column schedule once more, with 8 x 12 columns: CTAs of 160 threads (3 compute warps, halo
// warp, sync warp), four per SM at the full register count.  Entry points: launch_sweep_columns_ek12 etc. (see the head of
// sdfb_sweep_columns.cu for when each build is used).
#define SDFB_EK 12
#define SDFB_MINB_SMALL 4
#if defined(SDFB_MERGE_SYNC) && SDFB_MERGE_SYNC
#define SDFB_MINB_HI 5          // without the sync warp a CTA has 128 threads: five fit an SM at the full register count
#endif
#define SDFB_COLS_SUFFIX _ek12
#include "sdfb_sweep_columns.cu"
