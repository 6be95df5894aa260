// ep_peer.cu — expert parallelism over NVLink / NVSwitch peer memory (sm_100a, one process per GPU).
//
// Replaces FastMoE's fmoe_cuda.expert_exchange + global_scatter / global_gather (NCCL grouped send/recv with a host
// round trip for the counts; un-vendored upstream, the world_size > 1 mode of the layer imported at
// /root/reference/models/resMoE.py:6) and round 1's NCCL all_to_all_single on padded fixed slabs + four repack passes
// per layer.  Every rank owns one symmetric heap (cudaMalloc + CUDA IPC); the packed row buffers of the local experts
// live in it, and the dispatch / combine kernels of the OTHER ranks write / read their rows there directly:
//
//   forward    gate + scan -> [counts exchange + barrier + layout] -> dispatch: kept rows go straight into the owner's
//              packed segment (no send buffer, no receive buffer, no repack; only live rows cross NVLink)
//              -> barrier -> expert FFN on the packed buffer -> barrier -> combine reads the remote Y rows in place
//   backward   combine_bwd writes dY rows into the owner's buffer -> barrier -> expert FFN backward -> barrier ->
//              gate/dispatch backward gathers the remote dX rows in place
//
// Packed layout on the owner: expert-major, inside an expert segment the source ranks in rank order, inside a source
// its pairs in token order (the order of the single-GPU path restricted to that source) — a pure function of the
// kept counts, which every rank computes redundantly from the W x E table the ranks write into each other's heaps.
// The barrier is a flag exchange through the heaps (release store of a monotonically increasing epoch into every
// peer's flag slot, acquire spin on the own slots); epochs live in device memory, so a CUDA-graph replay keeps
// counting.  A spin that exceeds its bound sets `status` instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace moe {

namespace {

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int* peer_ints(const PeerRows& pr, int rk) {
    void* b = pr.base[0];
#pragma unroll
    for (int i = 1; i < kMaxPeers; ++i)
        if (rk == i) b = pr.base[i];
    return static_cast<int*>(b);
}

// All W ranks meet here.  Called by every thread of the (single) CTA after its own global writes; returns after
// every rank's writes that preceded its arrival are visible to this CTA.  threads 0..W-1 signal / wait.
__device__ __forceinline__ void cta_peer_barrier(const PeerRows& flags, int* epoch, int rank, int W, int* status, int* epoch_s) {
    __syncthreads();                       // the CTA's earlier writes happen-before the release below
    if (threadIdx.x == 0) *epoch_s = *epoch + 1;
    __syncthreads();
    const int ep = *epoch_s;
    if (static_cast<int>(threadIdx.x) < W) {
        __threadfence_system();
        st_release_sys(peer_ints(flags, threadIdx.x) + rank, ep);          // my arrival, in peer threadIdx.x's slot [rank]
        const int* mine = peer_ints(flags, rank) + threadIdx.x;            // peer threadIdx.x's arrival, in my slot
        unsigned spins = 0;
        while (ld_acquire_sys(mine) - ep < 0) {                            // epochs only grow (wrap-safe comparison)
            __nanosleep(64);
            if (++spins > (1u << 25)) { atomicExch(status, 1); break; }    // ~ seconds: report, never hang
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch = ep;
}

__global__ void __launch_bounds__(32)
ep_barrier_kernel(PeerRows flags, int* epoch, int rank, int W, int* status) {
    __shared__ int epoch_s;
    cta_peer_barrier(flags, epoch, rank, W, status, &epoch_s);
}

// kept[E]: this rank's kept pairs per GLOBAL expert (route_scan).  kept_all: every rank's [W][E] table.
__global__ void __launch_bounds__(1024)
ep_exchange_counts_kernel(const int* __restrict__ kept, PeerRows kept_all, PeerRows flags, int* epoch, int rank, int W, int El,
                          int rows_per_rank, int* __restrict__ dst_row, int* __restrict__ kept_local, int* __restrict__ seg_start,
                          int* __restrict__ tile_expert, int* __restrict__ num_mtiles, int max_mtiles, int* status) {
    extern __shared__ int smem_ep[];
    const int E = W * El;
    int* ka = smem_ep;                  // [W][E] all ranks' counts
    int* seg_all = ka + W * E;          // [W][El + 1] packed segment starts of every owner
    __shared__ int epoch_s;
    const int tid = threadIdx.x;
    for (int i = tid; i < W * E; i += 1024) {
        const int p = i / E, e = i - p * E;
        peer_ints(kept_all, p)[rank * E + e] = kept[e];
    }
    cta_peer_barrier(flags, epoch, rank, W, status, &epoch_s);
    const int* mine = peer_ints(kept_all, rank);
    for (int i = tid; i < W * E; i += 1024) ka[i] = __ldcv(mine + i);      // written by the peers: never from a stale L1 line
    __syncthreads();
    if (tid < W) {                      // thread o lays out owner o's buffer
        int start = 0;
        for (int l = 0; l < El; ++l) {
            seg_all[tid * (El + 1) + l] = start;
            int tot = 0;
            for (int s = 0; s < W; ++s) tot += ka[s * E + tid * El + l];
            start += (tot + MOE_ROW_ALIGN - 1) / MOE_ROW_ALIGN * MOE_ROW_ALIGN;
        }
        seg_all[tid * (El + 1) + El] = start;
        if (start > rows_per_rank) atomicExch(status, 2);                 // cannot happen with per-source capacities; checked anyway
    }
    __syncthreads();
    for (int e = tid; e < E; e += 1024) {
        const int o = e / El, l = e - o * El;
        int before = 0;
        for (int s = 0; s < rank; ++s) before += ka[s * E + e];
        dst_row[e] = o * rows_per_rank + seg_all[o * (El + 1) + l] + before;
    }
    for (int l = tid; l <= El; l += 1024) {
        seg_start[l] = seg_all[rank * (El + 1) + l];
        if (l < El) {
            int tot = 0;
            for (int s = 0; s < W; ++s) tot += ka[s * E + rank * El + l];
            kept_local[l] = tot;
        }
    }
    const int* seg_s = seg_all + rank * (El + 1);
    const int nm = seg_s[El] / MOE_ROW_ALIGN;
    if (tid == 0) *num_mtiles = nm;
    for (int m = tid; m < max_mtiles; m += 1024) {
        int e = -1;
        if (m < nm) {
            const int row = m * MOE_ROW_ALIGN;
            e = 0;
            while (e + 1 < El && seg_s[e + 1] <= row) ++e;
        }
        tile_expert[m] = e;
    }
}

}  // namespace

cudaError_t launch_ep_barrier(const PeerRows& flags, int* epoch, int rank, int W, int* status, cudaStream_t st) {
    ep_barrier_kernel<<<1, 32, 0, st>>>(flags, epoch, rank, W, status);
    return cudaGetLastError();
}

cudaError_t launch_ep_exchange_counts(const int* kept, const PeerRows& kept_all, const PeerRows& flags, int* epoch, int rank, int W,
                                      int El, int rows_per_rank, int* dst_row, int* kept_local, int* seg_start, int* tile_expert,
                                      int* num_mtiles, int max_mtiles, int* status, cudaStream_t st) {
    const size_t smem = (static_cast<size_t>(W) * W * El + static_cast<size_t>(W) * (El + 1)) * sizeof(int);
    ep_exchange_counts_kernel<<<1, 1024, smem, st>>>(kept, kept_all, flags, epoch, rank, W, El, rows_per_rank, dst_row, kept_local,
                                                     seg_start, tile_expert, num_mtiles, max_mtiles, status);
    return cudaGetLastError();
}

}  // namespace moe
