// ipc_probe.cu -- can two PROCESSES on one box (one GPU each, or both on one GPU) share device
// memory through CUDA IPC, signal each other with flags written into the peer's memory from inside
// a kernel, and read 136-byte block records straight out of the peer's memory?  These are the three
// things the peer-memory exchange of the sharded map build relies on (vmap_shard_host.inl).
//   build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/ipc_probe tools/ipc_probe.cu
//   run:   tools/ipc_probe            (prints one JSON line per rank; exit code 0 = all checks passed)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "rank %d: %s: %s\n", g_rank, #x, cudaGetErrorString(e_)); exit(2); } } while (0)
static int g_rank = 0;

constexpr size_t kFlagBytes = 4096;
constexpr uint32_t kRecWords = 34;            // 136-byte record
constexpr uint32_t kRecords = 1u << 19;       // 71 MB

__device__ __forceinline__ uint32_t ld_flag(const volatile uint32_t* p) { return *p; }

// flag ping-pong: rank 0 starts.  mine = flags in MY memory (written by the peer), theirs = the
// peer's flag area.  Returns through `out` the number of spins that timed out (0 = fine).
__global__ void k_pingpong(volatile uint32_t* mine, volatile uint32_t* theirs, int rank, int iters, uint32_t* out) {
  uint32_t bad = 0;
  for (int i = 1; i <= iters; i++) {
    if (rank == 0) {
      __threadfence_system();
      theirs[0] = (uint32_t)i;
      long long t0 = clock64();
      while (ld_flag(mine + 1) != (uint32_t)i) if (clock64() - t0 > 4000000000ll) { bad++; break; }
    } else {
      long long t0 = clock64();
      while (ld_flag(mine + 0) != (uint32_t)i) if (clock64() - t0 > 4000000000ll) { bad++; break; }
      __threadfence_system();
      theirs[1] = (uint32_t)i;
    }
  }
  *out = bad;
}

__global__ void k_fill(uint32_t* rec, uint32_t n, uint32_t seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n * kRecWords; i += (size_t)gridDim.x * blockDim.x)
    rec[i] = (uint32_t)(i * 2654435761u) ^ seed;
}

// tell the peer "my records are ready" / wait for the peer's: one thread
__global__ void k_signal(volatile uint32_t* theirs, int slot, uint32_t v) { __threadfence_system(); theirs[slot] = v; }
__global__ void k_wait(volatile uint32_t* mine, int slot, uint32_t v, uint32_t* timeout) {
  long long t0 = clock64();
  while (ld_flag(mine + slot) < v) { __nanosleep(200); if (clock64() - t0 > 8000000000ll) { *timeout = 1; return; } }
}

// one warp per record: lanes 0..33 would be ideal; here 32 lanes read 32 words + lanes 0,1 the tail
__global__ void k_read_records(const uint32_t* __restrict__ rec, uint32_t n, unsigned long long* sum) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31u;
  unsigned long long acc = 0;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    const uint32_t* p = rec + (size_t)r * kRecWords;
    acc += p[lane];
    if (lane < 2) acc += p[32 + lane];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if (lane == 0) atomicAdd(sum, acc);
}

static unsigned long long expect_sum(uint32_t seed) {
  unsigned long long s = 0;
  for (size_t i = 0; i < (size_t)kRecords * kRecWords; i++) s += (uint32_t)((uint32_t)(i * 2654435761u) ^ seed);
  return s;
}

int main() {
  int to_child[2], to_parent[2];
  if (pipe(to_child) || pipe(to_parent)) return 3;
  const pid_t pid = fork();
  g_rank = pid == 0 ? 1 : 0;
  const int rfd = g_rank ? to_child[0] : to_parent[0], wfd = g_rank ? to_parent[1] : to_child[1];
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const int dev = ndev > 1 ? g_rank : 0;
  CK(cudaSetDevice(dev));
  uint8_t* base = nullptr;
  const size_t bytes = kFlagBytes + (size_t)kRecords * kRecWords * 4;
  CK(cudaMalloc(&base, bytes));
  CK(cudaMemset(base, 0, kFlagBytes));
  cudaIpcMemHandle_t mine_h, peer_h;
  CK(cudaIpcGetMemHandle(&mine_h, base));
  if (write(wfd, &mine_h, sizeof(mine_h)) != (ssize_t)sizeof(mine_h)) return 3;
  if (read(rfd, &peer_h, sizeof(peer_h)) != (ssize_t)sizeof(peer_h)) return 3;
  uint8_t* peer = nullptr;
  CK(cudaIpcOpenMemHandle((void**)&peer, peer_h, cudaIpcMemLazyEnablePeerAccess));
  volatile uint32_t* my_flags = (volatile uint32_t*)base;
  volatile uint32_t* peer_flags = (volatile uint32_t*)peer;
  uint32_t* my_rec = (uint32_t*)(base + kFlagBytes);
  const uint32_t* peer_rec = (const uint32_t*)(peer + kFlagBytes);
  uint32_t* d_out = nullptr;
  unsigned long long* d_sum = nullptr;
  CK(cudaMalloc(&d_out, 16));
  CK(cudaMalloc(&d_sum, 8));
  CK(cudaMemset(d_out, 0, 16));
  CK(cudaMemset(d_sum, 0, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  // host handshake so both kernels are about to run
  char c = 'x';
  if (write(wfd, &c, 1) != 1 || read(rfd, &c, 1) != 1) return 3;

  const int iters = ndev > 1 ? 2000 : 2;   // (one GPU: the two processes time-slice, a ping-pong costs two slices)
  CK(cudaEventRecord(e0));
  k_pingpong<<<1, 1>>>(my_flags, peer_flags, g_rank, iters, d_out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms_pp = 0;
  CK(cudaEventElapsedTime(&ms_pp, e0, e1));
  uint32_t bad = 0;
  CK(cudaMemcpy(&bad, d_out, 4, cudaMemcpyDeviceToHost));

  // records: fill mine, signal, wait for the peer's, read the peer's
  k_fill<<<148 * 8, 256>>>(my_rec, kRecords, 0x1234u + g_rank);
  k_signal<<<1, 1>>>(peer_flags, 2 + g_rank, 1u);
  k_wait<<<1, 1>>>(my_flags, 2 + (g_rank ^ 1), 1u, d_out + 1);
  float best = 1e30f;
  unsigned long long sum = 0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaMemset(d_sum, 0, 8));
    CK(cudaEventRecord(e0));
    k_read_records<<<148 * 8, 256>>>(peer_rec, kRecords, d_sum);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
    CK(cudaMemcpy(&sum, d_sum, 8, cudaMemcpyDeviceToHost));
  }
  uint32_t timeout = 0;
  CK(cudaMemcpy(&timeout, d_out + 1, 4, cudaMemcpyDeviceToHost));
  const bool sum_ok = sum == expect_sum(0x1234u + (g_rank ^ 1));
  const double gbs = (double)kRecords * kRecWords * 4 / (best * 1e-3) / 1e9;
  printf("{\"rank\": %d, \"device\": %d, \"n_devices\": %d, \"ipc_open\": true, \"pingpong_timeouts\": %u, "
         "\"pingpong_round_trip_us\": %.2f, \"wait_timeout\": %u, \"peer_read_ok\": %s, \"peer_read_gbs\": %.1f}\n",
         g_rank, dev, ndev, bad, 1e3 * ms_pp / iters, timeout, sum_ok ? "true" : "false", gbs);
  fflush(stdout);
  // keep the memory alive until the peer is done too
  if (write(wfd, &c, 1) != 1 || read(rfd, &c, 1) != 1) return 3;
  CK(cudaIpcCloseMemHandle(peer));
  CK(cudaFree(base));
  const int ok = (!bad && !timeout && sum_ok) ? 0 : 1;
  if (g_rank == 0) {
    int st = 0;
    waitpid(pid, &st, 0);
    return ok || !WIFEXITED(st) || WEXITSTATUS(st);
  }
  return ok;
}
