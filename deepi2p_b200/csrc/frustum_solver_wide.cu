// Second build of frustum_solver.cu with 128-thread CTAs, in its own namespace (dib_w128), reached only through
// the primary TU's C ABI for batches with few problems (DIB_WIDE_BELOW, experimental, default off).
// The whole file is re-used as is: `dib` is renamed by the preprocessor, the C ABI and the error buffer are
// compiled out by DIB_WIDE_TU, and everything shared with the other TUs (common.cuh) keeps its real name
// through the alias below.
#include "common.cuh"

namespace dibcore = dib;        // spelled differently so that the rename below leaves it alone

#undef DIB_THREADS              // a tuning build may pass -DDIB_THREADS=... for the primary TU
#undef DIB_MINBLOCKS4
#undef DIB_MINBLOCKS6
#define DIB_WIDE_TU 1
#define DIB_THREADS 128
#define DIB_MINBLOCKS4 5
#define DIB_MINBLOCKS6 4
#define dib dib_w128
namespace dib {
using namespace dibcore;        // device helpers of common.cuh
using dibcore::set_error;       // host error plumbing (defined in the primary TU)
}  // namespace dib_w128
#include "frustum_solver.cu"
