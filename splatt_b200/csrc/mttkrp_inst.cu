// Instantiation unit: compiled once per mode count (-DSPB200_INST_N=2..8) so the
// 6 x 4 x 3 kernel variants build in parallel.
#include "mttkrp_kernels.cuh"

#ifndef SPB200_INST_N
#error "compile with -DSPB200_INST_N=<nmodes>"
#endif

int spb200_root_batch();

namespace spb200 {

template <int N, int L, int KIND, int BATCH, bool KT = false, bool MC = false, int MINB = 0>
static int launch_variant(const MttkrpArgs & args, int num_sms, cudaStream_t stream) {
  auto kern = [] {
    if constexpr (MINB == 0) return mttkrp_stream_kernel<N, L, KIND, BATCH, KT, MC>;
    else return mttkrp_stream_kernel<N, L, KIND, BATCH, KT, MC, MINB>;
  }();
  static int occ = 0;   // per-variant, set once
  if (occ == 0) {
    SPB200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kSmemBytes)));
    int o = 0;
    SPB200_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, kSmemBytes));
    occ = o > 0 ? o : 1;
  }
  constexpr int G = 32 / L;
  // never launch more groups than there are chunks
  unsigned long long want = (static_cast<unsigned long long>(args.nchunks) + kWarps * G - 1) / (kWarps * G);
  unsigned long long grid = static_cast<unsigned long long>(num_sms) * occ;
  if (grid > want) grid = want;
  if (grid == 0) return SPLATT_SUCCESS;
  kern<<<static_cast<unsigned>(grid), kThreads, kSmemBytes, stream>>>(args);
  ++g_spb200_launches;
  SPB200_CUDA_OK(cudaGetLastError());
  return SPLATT_SUCCESS;
}

template <int N, int L>
static int launch_kind(int kind, const MttkrpArgs & args, int num_sms, cudaStream_t stream) {
  switch (kind) {
    case SPB200_KIND_ROOT:
      if (args.multicast) return launch_variant<N, L, SPB200_KIND_ROOT, 4, false, true>(args, num_sms, stream);
      if (spb200_root_batch() >= 8) return launch_variant<N, L, SPB200_KIND_ROOT, 8>(args, num_sms, stream);
      // deeper trees hold a third gathered row per record: two-record batches keep the
      // kernel at 80 registers / 3 CTAs per SM (measured 1014 vs 1052 us on config 3)
      if (spb200_root_batch() == 2 || (spb200_root_batch() == 0 && N >= 4))
        return launch_variant<N, L, SPB200_KIND_ROOT, 2>(args, num_sms, stream);
      if (args.ktiled) return launch_variant<N, L, SPB200_KIND_ROOT, 4, true>(args, num_sms, stream);
      return launch_variant<N, L, SPB200_KIND_ROOT, 4>(args, num_sms, stream);
    case SPB200_KIND_INTL: return launch_variant<N, L, SPB200_KIND_INTL, 4>(args, num_sms, stream);
    default:               return launch_variant<N, L, SPB200_KIND_LEAF, 4>(args, num_sms, stream);
  }
}

#define SPB200_CAT_(a, b) a##b
#define SPB200_CAT(a, b) SPB200_CAT_(a, b)

// ncols: active (even) columns of this launch, <= 64.
int SPB200_CAT(launch_n, SPB200_INST_N)(int kind, const MttkrpArgs & args, int num_sms,
                                         cudaStream_t stream) {
  constexpr int N = SPB200_INST_N;
  if (args.ncols <= 8)  return launch_kind<N, 4>(kind, args, num_sms, stream);
  if (args.ncols <= 16) return launch_kind<N, 8>(kind, args, num_sms, stream);
  if (args.ncols <= 32) return launch_kind<N, 16>(kind, args, num_sms, stream);
  return launch_kind<N, 32>(kind, args, num_sms, stream);
}

}  // namespace spb200
