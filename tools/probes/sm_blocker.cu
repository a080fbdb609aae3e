// Occupies `n` SMs completely (one 1024-thread CTA with 227 KB of shared memory each) for `ns` nanoseconds.
// tools/probe_overlap.py launches it beside inverse_fill to see how the fill's time depends on the SMs it gets.
//   nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tools/_build/libsm_blocker.so tools/probes/sm_blocker.cu
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void __launch_bounds__(1024) blocker_kernel(unsigned long long ns, int* sink) {
  extern __shared__ int sm[];
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    __nanosleep(200);
  } while (t - t0 < ns);
  if (ns == 0x7fffffffffffffffull) sink[0] = sm[threadIdx.x];
}
extern "C" int sm_blocker(int n, unsigned long long ns, int* sink, void* stream) {
  const int smem = 227 * 1024;   // the per-CTA maximum: with the 1 KB the system reserves per CTA nothing else fits on the SM
  cudaFuncSetAttribute(blocker_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  blocker_kernel<<<n, 1024, smem, static_cast<cudaStream_t>(stream)>>>(ns, sink);
  return static_cast<int>(cudaGetLastError());
}
