// randrec.cu - microbenchmark: what can B200 sustain when every LANE reads (and partly rewrites) its
// own randomly placed record of R bytes in a multi-GB arena? This is the access pattern of the
// device-resident search (one lane = one tree, one 288-byte node record per level), used to put a
// realistic ceiling next to the streaming-copy peak in MEASURED_PEAKS.json.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o randrec randrec.cu && ./randrec [GB]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// R = record bytes (multiple of 16), W = bytes written back per visit, ILP independent records per lane
template <int R, int W, int ILP>
__global__ void __launch_bounds__(256) visit(uint4* arena, uint64_t n_records, uint32_t visits, uint64_t seed,
                                            unsigned long long* sink) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t acc = 0;
  for (uint32_t v = 0; v < visits; ++v) {
    uint4 buf[ILP][R / 16];
    uint64_t rec[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      rec[k] = mix(seed + tid * 0x9E3779B97F4A7C15ULL + (uint64_t)v * ILP + k + acc * 0) % n_records;
      const uint4* p = arena + rec[k] * (R / 16);
#pragma unroll
      for (int i = 0; i < R / 16; ++i) buf[k][i] = p[i];
    }
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      uint4* p = arena + rec[k] * (R / 16);
#pragma unroll
      for (int i = 0; i < R / 16; ++i) acc += buf[k][i].x;
#pragma unroll
      for (int i = 0; i < W / 16; ++i) { uint4 w = buf[k][i]; w.x += 1; p[i] = w; }
    }
  }
  if (acc == 0x123456789ULL) atomicAdd(sink, acc);
}

// Private-region variant: lane t only touches records inside its own region of `region_recs` records
// (the search: one lane = one tree arena of 2.88 MB = two 2-MB pages), so the translations a lane needs
// stay the same for the whole launch.
template <int R, int W, int ILP>
__global__ void __launch_bounds__(64) visit_private(uint4* arena, uint64_t region_recs, uint32_t lanes, uint32_t visits,
                                                   uint64_t seed, unsigned long long* sink) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (tid >= lanes) return;
  uint4* base = arena + tid * region_recs * (R / 16);
  uint64_t acc = 0;
  for (uint32_t v = 0; v < visits; ++v) {
    uint4 buf[ILP][R / 16];
    uint64_t rec[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      rec[k] = mix(seed + tid * 0x9E3779B97F4A7C15ULL + (uint64_t)v * ILP + k) % region_recs;
      const uint4* p = base + rec[k] * (R / 16);
#pragma unroll
      for (int i = 0; i < R / 16; ++i) buf[k][i] = p[i];
    }
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      uint4* p = base + rec[k] * (R / 16);
#pragma unroll
      for (int i = 0; i < R / 16; ++i) acc += buf[k][i].x;
#pragma unroll
      for (int i = 0; i < W / 16; ++i) { uint4 w = buf[k][i]; w.x += 1; p[i] = w; }
    }
  }
  if (acc == 0x123456789ULL) atomicAdd(sink, acc);
}

template <int R, int W, int ILP>
static void run_private(uint4* arena, uint32_t lanes, uint64_t region_recs, unsigned long long* sink) {
  const uint32_t visits = 2000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = (lanes + 63) / 64;
  visit_private<R, W, ILP><<<blocks, 64>>>(arena, region_recs, lanes, 16, 1, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  visit_private<R, W, ILP><<<blocks, 64>>>(arena, region_recs, lanes, visits, 7, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const double n = (double)lanes * visits * ILP;
  printf("private R=%3d W=%3d ILP=%d lanes=%6u region=%7.2f MB footprint=%6.1f GB : %.3f Grec/s  total %.0f GB/s\n", R, W, ILP,
         lanes, region_recs * R / 1048576.0, (double)lanes * region_recs * R / 1073741824.0, n / ms * 1e-6,
         n * (R + W) / ms * 1e-6);
}

template <int R, int W, int ILP>
static void run(uint4* arena, size_t bytes, int blocks, unsigned long long* sink) {
  const uint64_t n_records = bytes / R;
  const uint32_t visits = 64;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  visit<R, W, ILP><<<blocks, 256>>>(arena, n_records, 4, 1, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  visit<R, W, ILP><<<blocks, 256>>>(arena, n_records, visits, 7, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const double n = (double)blocks * 256 * visits * ILP;
  printf("R=%3d W=%3d ILP=%d blocks=%5d : %.3f Grec/s  read %.0f GB/s  write %.0f GB/s  total %.0f GB/s\n", R, W, ILP,
         blocks, n / ms * 1e-6, n * R / ms * 1e-6, n * W / ms * 1e-6, n * (R + W) / ms * 1e-6);
}

int main(int argc, char** argv) {
  const size_t gb = argc > 1 ? atoi(argv[1]) : 64;
  const size_t bytes = gb << 30;
  uint4* arena; unsigned long long* sink;
  if (cudaMalloc(&arena, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&sink, 8);
  cudaMemset(arena, 1, bytes);
  if (argc > 2) {   // private-region mode
    for (uint32_t lanes : {16384u, 32768u}) {
      for (uint64_t recs : {512ull, 2048ull, 10001ull}) {
        if ((double)lanes * recs * 288 > (double)bytes) continue;
        run_private<288, 192, 1>(arena, lanes, recs, sink);
        run_private<288, 192, 4>(arena, lanes, recs, sink);
      }
    }
    return 0;
  }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int bps : {1, 2, 4, 8}) {
    const int blocks = sms * bps;
    run<32, 32, 1>(arena, bytes, blocks, sink);
    run<64, 64, 1>(arena, bytes, blocks, sink);
    run<256, 192, 1>(arena, bytes, blocks, sink);
    run<288, 192, 1>(arena, bytes, blocks, sink);
    run<256, 192, 2>(arena, bytes, blocks, sink);
    run<256, 0, 1>(arena, bytes, blocks, sink);
  }
  return 0;
}
