// sdfb_sweep_columns.cu -- placeholder until the pipelined column schedule lands: routes to the
// per-level schedule so the ABI is complete.
#include "sdfb_kernels.cuh"
namespace sdfb {
size_t sweep_columns_progress_words(const Grid &) { return 0; }
int launch_sweep_columns(uint64_t *cells, const TriRec *rec, const Grid &g, int sweep_index,
                         unsigned long long *changed, uint32_t *, cudaStream_t st)
{
    return launch_sweep_levels(cells, rec, g, sweep_index, changed, st);
}
}  // namespace sdfb
