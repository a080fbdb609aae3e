// standalone check of the TMA tile-load helpers in csrc/tma.cuh (nvcc -arch=sm_100a ... ; run on the B200 box)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../foveated-instance-segmentation_b200/csrc/tma.cuh"  // nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -o tools/_build/tma_load_test tools/probes/tma_load_test.cu
namespace fovea { void set_error(const char* fmt, ...) { printf("error: %s\n", fmt); } }
using namespace fovea;
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int x, int y, int z, int step) {
  __shared__ __align__(128) float box[32][64];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (step >= 1 && threadIdx.x == 0) mbar_expect_tx(&bar, 32 * 64 * 4);
  if (step >= 2 && threadIdx.x == 0) tma_load_tile(&tmap, &box[0][0], &bar, x, y, z);
  if (step >= 3) mbar_wait(&bar, 0);
  __syncthreads();
  if (threadIdx.x < 8) out[threadIdx.x] = step >= 3 ? box[threadIdx.x / 4][threadIdx.x % 4] : 1.f;
}
int main(int argc, char** argv) {
  const int cx = argc > 1 ? atoi(argv[1]) : -1, cy = argc > 2 ? atoi(argv[2]) : -1, cz = argc > 3 ? atoi(argv[3]) : 1;
  const int P = 3, H = 64, W = 64;
  std::vector<float> h(P * H * W);
  for (int i = 0; i < P * H * W; ++i) h[i] = float(i % 1000);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 64);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap map;
  if (make_plane_load_map(&map, d, P, H, W, 64, 32)) return 1;
  for (int step = 0; step <= 3; ++step) {
    k<<<1, 256>>>(map, o, cx, cy, cz, step);
    cudaError_t e = cudaDeviceSynchronize();
    float r[8]; cudaMemcpy(r, o, 32, cudaMemcpyDeviceToHost);
    printf("step %d: %s  out = %g %g %g %g | %g %g %g %g\n", step, cudaGetErrorString(e), r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
