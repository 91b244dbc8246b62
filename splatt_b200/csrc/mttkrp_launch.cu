// Host-side dispatch of one MTTKRP over a fiber stream.
#include "common.h"
#include <cstdlib>

unsigned long long g_spb200_launches = 0;

// Tuning knob (experiments): records per gather batch of the root kernel.
int spb200_root_batch() {
  static int v = -1;
  if (v < 0) {
    const char * e = getenv("SPLATT_B200_BATCH");
    v = e ? atoi(e) : 0;   // 0 = per-kernel default (4)
    if (v != 2 && v != 8) v = 4;
  }
  return v;
}

namespace spb200 {
int launch_n2(int, const MttkrpArgs &, int, cudaStream_t);
int launch_n3(int, const MttkrpArgs &, int, cudaStream_t);
int launch_n4(int, const MttkrpArgs &, int, cudaStream_t);
int launch_n5(int, const MttkrpArgs &, int, cudaStream_t);
int launch_n6(int, const MttkrpArgs &, int, cudaStream_t);
int launch_n7(int, const MttkrpArgs &, int, cudaStream_t);
int launch_n8(int, const MttkrpArgs &, int, cudaStream_t);
}  // namespace spb200

int spb200_root_minb() {
  static int v = -1;
  if (v < 0) {
    const char * e = getenv("SPLATT_B200_MINB");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static int num_sms_of_current_device() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int spb200_launch_mttkrp(const FiberStream & s, int kind, int outdepth, int ncolumns, int ldm,
                         const double * const * d_mats_by_mode, double * d_out,
                         uint64_t out_rows, cudaStream_t stream, bool multicast_out) {
  const int N = s.nmodes;
  if (N < 2 || N > SPB200_MAXN) {
    fprintf(stderr, "SPLATT: MTTKRP supports 2..%d modes (got %d)\n", SPB200_MAXN, N);
    return SPLATT_ERROR_BADINPUT;
  }
  if (ncolumns <= 0 || ldm < ncolumns + (ncolumns & 1) || (ldm & 1)) {
    fprintf(stderr, "SPLATT: bad ncolumns/ldm (%d/%d): ldm must be even and >= ncolumns\n",
            ncolumns, ldm);
    return SPLATT_ERROR_BADINPUT;
  }
  if (multicast_out) {
    // the caller zeroed every GPU's buffer and synchronised the group beforehand
    if (kind != SPB200_KIND_ROOT) {
      fprintf(stderr, "SPLATT: multicast output needs a root-oriented stream (ALLROOT layout)\n");
      return SPLATT_ERROR_BADINPUT;
    }
  } else {
    SPB200_CUDA_OK(cudaMemsetAsync(d_out, 0, sizeof(double) * out_rows * ldm, stream));
  }
  if (s.nrec == 0) return SPLATT_SUCCESS;

  MttkrpArgs a;
  a.rec = s.rec;
  for (int l = 0; l < SPB200_MAXN - 2; ++l) a.up[l] = (l <= N - 3) ? s.up[l] : nullptr;   // none for N = 2
  a.desc = s.desc;
  for (int l = 0; l < SPB200_MAXN; ++l) a.mats[l] = (l < N) ? d_mats_by_mode[s.perm[l]] : nullptr;
  a.out      = d_out;
  a.nrec     = s.nrec;
  a.nchunks  = static_cast<unsigned int>(s.nchunks);
  a.ldm      = ldm;
  a.outdepth = outdepth;
  a.ktiled   = s.ktile_rows ? 1 : 0;
  a.multicast = multicast_out ? 1 : 0;

  const int rpad    = ncolumns + (ncolumns & 1);
  const int num_sms = num_sms_of_current_device();
  for (int c0 = 0; c0 < rpad; c0 += 64) {
    a.col0  = c0;
    a.ncols = (rpad - c0 < 64) ? (rpad - c0) : 64;
    int rc;
    switch (N) {
      case 2:  rc = spb200::launch_n2(kind, a, num_sms, stream); break;
      case 3:  rc = spb200::launch_n3(kind, a, num_sms, stream); break;
      case 4:  rc = spb200::launch_n4(kind, a, num_sms, stream); break;
      case 5:  rc = spb200::launch_n5(kind, a, num_sms, stream); break;
      case 6:  rc = spb200::launch_n6(kind, a, num_sms, stream); break;
      case 7:  rc = spb200::launch_n7(kind, a, num_sms, stream); break;
      default: rc = spb200::launch_n8(kind, a, num_sms, stream); break;
    }
    if (rc != SPLATT_SUCCESS) return rc;
  }
  return SPLATT_SUCCESS;
}
