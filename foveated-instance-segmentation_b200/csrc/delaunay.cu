// Device Delaunay triangulation (placeholder until the kernel lands; the host-Qhull parity mode does not use it).
#include "common.cuh"

using namespace fovea;

extern "C" int64_t fovea_delaunay_workspace_bytes(int B, int cap) {
  (void)B; (void)cap;
  return 16;
}

extern "C" int fovea_delaunay(const int32_t* pts, const int32_t* npts, int B, int cap, int tcap, uint16_t* tris,
                              uint16_t* nbrs, int32_t* ntri, void* workspace, fovea_stream_t stream) {
  (void)pts; (void)npts; (void)B; (void)cap; (void)tcap; (void)tris; (void)nbrs; (void)ntri; (void)workspace; (void)stream;
  set_error("fovea_delaunay: device triangulation is not built into this library");
  return FOVEA_ERR_ARG;
}
