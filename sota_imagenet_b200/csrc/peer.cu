// One-shot all-reduce of small fp32 vectors over NVLink peer memory (SyncBN statistics).
//
// The data-parallel step all-reduces one [2][C] (or [4][C]) fp32 vector per BatchNorm and pass:
// 106 sequentially dependent reductions of <= 32 KB per ResNet-50 step (reference: SyncBatchNorm
// under DistributedDataParallel, train.py:113-114).  Through NCCL each costs a full collective
// latency on the critical path.  Here every rank owns a "mailbox" in device memory that its peers
// map with CUDA IPC; one kernel per reduction
//   1. PUSHES the local vector into slot [rank] of every peer's mailbox as 8-byte {value, epoch}
//      words (NVLink stores, "LL" protocol: the epoch travels with the data),
//   2. polls its own mailbox until every word of every rank carries this call's epoch and sums
//      the W slots in rank order -> bitwise identical result on all ranks.
// No collective library call, no host synchronisation; capturable in a CUDA graph (all addresses
// are fixed; the epoch lives in device memory and is bumped by the kernel itself).  Mailboxes are
// double-buffered by epoch parity, so a rank that is one call ahead can never overwrite data a
// slower peer is still summing.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"
#include "../../include/sib200.h"

namespace sib {

constexpr int kPeerMaxWorld = 8;

struct PeerTable {
  unsigned long long* mailbox[kPeerMaxWorld];   // [2 parities][capacity words], per rank
  unsigned* unused[kPeerMaxWorld];
};

// "LL" protocol (as in NCCL's low-latency path): every float travels as one 8-byte word
// {value bits, epoch}; an aligned 8-byte store is delivered atomically over NVLink, so the
// receiver simply polls each word until it carries this call's epoch -- no fence, no separate
// flag, one NVLink hop of latency.
__device__ __forceinline__ void st_ll(unsigned long long* p, float v, unsigned epoch) {
  const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

__global__ void __launch_bounds__(512)
peer_allreduce_kernel(float* __restrict__ data, int n, const PeerTable* __restrict__ tab,
                      long slot_off, long parity_stride, int slot, int num_slots,
                      unsigned* __restrict__ epochs, int rank, int world) {
  __shared__ unsigned s_epoch;
  __shared__ PeerTable t;
  if (threadIdx.x == 0) {
    s_epoch = epochs[slot] + 1;
    t = *tab;
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const long base = (long)(epoch & 1u) * parity_stride + slot_off;
  // 1. push this rank's vector into slot [rank] of every mailbox (its own included)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = data[i];
    for (int p = 0; p < world; ++p) st_ll(t.mailbox[p] + base + (long)rank * n + i, v, epoch);
  }
  // 2. poll the own mailbox word by word and sum in rank order (bitwise identical everywhere)
  const unsigned long long* mb = t.mailbox[rank] + base;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < world; ++q) {
      unsigned long long w = ld_ll(mb + (long)q * n + i);
      unsigned spins = 0;
      while ((unsigned)(w >> 32) != epoch) {
        if (++spins > (1u << 26)) {
          printf("sib: peer all-reduce timeout rank %d slot %d waiting for rank %d\n", rank, slot, q);
          __trap();
        }
        w = ld_ll(mb + (long)q * n + i);
      }
      s += __uint_as_float((unsigned)w);
    }
    data[i] = s;
  }
  if (threadIdx.x == 0) epochs[slot] = epoch;
}

}  // namespace sib

using namespace sib;

extern "C" int sib_ipc_alloc(unsigned long long bytes, void** ptr, unsigned char* handle64) {
  SIB_CHECK(sizeof(cudaIpcMemHandle_t) == 64, "unexpected cudaIpcMemHandle_t size");
  void* p = nullptr;
  SIB_CUDA(cudaMalloc(&p, bytes));
  SIB_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  SIB_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

extern "C" int sib_ipc_open(const unsigned char* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  SIB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return 0;
}

extern "C" int sib_ipc_close(void* ptr) {
  SIB_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int sib_ipc_free(void* ptr) {
  SIB_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int sib_peer_allreduce(float* data, int n, const void* table_dev, long slot_off,
                                  long parity_stride, int slot, int num_slots, void* epochs_dev,
                                  int rank, int world, void* stream) {
  SIB_CHECK(world >= 1 && world <= kPeerMaxWorld, "peer all-reduce: world size %d unsupported", world);
  SIB_CHECK(n > 0 && (reinterpret_cast<uintptr_t>(data) & 3) == 0, "peer all-reduce: bad vector");
  SIB_CHECK(slot >= 0 && slot < num_slots, "peer all-reduce: slot %d out of range", slot);
  peer_allreduce_kernel<<<1, n >= 2048 ? 512 : 256, 0, static_cast<cudaStream_t>(stream)>>>(
      data, n, static_cast<const PeerTable*>(table_dev), slot_off, parity_stride, slot, num_slots,
      static_cast<unsigned*>(epochs_dev), rank, world);
  SIB_LAUNCH_CHECK();
  return 0;
}
