// Grouped bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), hand-written for sm_100a.
//
//   C[g] = act( A[g] · B[g]^T + bias[g] ),   A: [G, M, K]  B: [G, N, K]  (both K-contiguous),  C: [G, M, N]  bf16, fp32 accumulate
//
// This is the MoE expert FFN (the one GEMM-shaped hot op of the framework; the reference runs a python loop of cuBLAS
// linears, bagua/torch_api/model_parallel/moe/experts.py:31-41). Persistent kernel, one CTA per SM walking 128 x BN
// output tiles (BN = 256, or 128 when N is not a multiple of 256):
//   warp 0      : TMA producer  — cp.async.bulk.tensor (128B-swizzled 128x64 / BNx64 tiles of A and B) into a 4/6-stage smem ring
//   warp 1      : MMA issuer    — one elected thread issues tcgen05.mma (M128 N{128,256} K16, kind::f16) per 32-byte K slice,
//                                  accumulating in one of TWO TMEM accumulators; tcgen05.commit releases smem stages and
//                                  signals the epilogue, which drains accumulator i while the MMAs fill accumulator i^1
//   warps 2..5  : epilogue      — tcgen05.ld the fp32 accumulator (32 lanes x 32 columns per instruction), + bias, GELU,
//                                  convert to bf16, 64-byte stores — to local memory, or (PEER variant) straight into the
//                                  symmetric buffer of the GPU that owns the token (the MoE "combine" all-to-all happens in
//                                  the epilogue, tile by tile, overlapped with the MMAs of the next tile)
// Synchronisation is mbarrier-only (full/empty per stage + accumulator full/empty). All waits are bounded: a wedged
// pipeline traps instead of hanging the GPU.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "kernels.h"

namespace bagua {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;
constexpr uint32_t kStageBytesA = BM * BK * 2;

// Persistent kernel, tile BM x BN with BN in {128, 256}. BN = 256 keeps the tensor pipe fed from shared memory: one
// M128 N256 K16 MMA takes 128 cycles and reads 4 KB of A + 8 KB of B = 96 B/cycle (the smem port moves 128 B/cycle; at
// N128 it would be exactly 128 B/cycle, i.e. smem-bound). Two accumulators of BN columns each double-buffer TMEM so
// the epilogue of tile i overlaps the MMAs of tile i+1.
template <int BN>
struct GemmCfg {
    static constexpr int STAGES = BN == 256 ? 4 : 6;
    static constexpr uint32_t kStageBytesB = BN * BK * 2;
    static constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512 (power of two)
};

template <int BN>
struct __align__(1024) GemmSmem {
    uint8_t a[GemmCfg<BN>::STAGES][kStageBytesA];
    uint8_t b[GemmCfg<BN>::STAGES][GemmCfg<BN>::kStageBytesB];
    uint64_t full[GemmCfg<BN>::STAGES];
    uint64_t empty[GemmCfg<BN>::STAGES];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();  // seconds of waiting: the pipeline is wedged — fail loudly, never hang
    }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// K-major operand tile in smem with the 128-byte swizzle written by TMA: rows of 128 B, 8-row atoms of 1024 B.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::f16: D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1, K-major A/B, N>>3 [17,23), M>>4 [24,29)
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_instr_desc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float gelu_tanh(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float u = k0 * (x + k1 * x * x * x);
    return 0.5f * x * (1.f + tanhf(u));
}

// tile id → (group, m block, n block): consecutive ids (= concurrently running CTAs) share the B tile and walk along M
struct TileCoord {
    int g, m0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, int tiles_m, int tiles_n, int BN) {
    TileCoord c;
    c.m0 = (tile % tiles_m) * BM;
    const int r = tile / tiles_m;
    c.n0 = (r % tiles_n) * BN;
    c.g = r / tiles_n;
    return c;
}

// PEER variant: where the epilogue stores go. Rows of A are ordered [source rank][capacity slot]; the tile of rows
// [m0, m0+128) of expert g belongs to source rank m0 / cap and lands in THAT rank's buffer, laid out
// [owner rank][expert][slot][N] so the receiver finds "what expert e on rank r produced for my slot c".
struct GemmPeerOut {
    char* ptr[kMaxPeers];
    size_t off;
    int rank;    // this (owner) rank
    int cap;     // capacity slots per (source rank, expert); multiple of BM
    int groups;  // local experts
};

template <int BN, bool PEER>
__global__ void __launch_bounds__(kGemmThreads, 1)
    grouped_gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, __nv_bfloat16* __restrict__ C,
                           const float* __restrict__ bias, int M, int N, int K, int num_tiles, int tiles_m, int tiles_n, int act,
                           const GemmPeerOut peer) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    GemmSmem<BN>& sm = *reinterpret_cast<GemmSmem<BN>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = K / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.acc_full[i], 1);
            mbar_init(&sm.acc_empty[i], 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // whole warp: TMEM allocation (the same warp frees it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(Cfg::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            uint32_t it = 0;  // running stage counter across tiles
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord(tile, tiles_m, tiles_n, BN);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&sm.empty[s], ((it / STAGES) & 1) ^ 1);  // fresh barrier: waiting on parity 1 passes immediately
                    mbar_expect_tx(&sm.full[s], kStageBytesA + Cfg::kStageBytesB);
                    tma_load_3d(sm.a[s], &tm_a, &sm.full[s], kb * BK, tc.m0, tc.g);
                    tma_load_3d(sm.b[s], &tm_b, &sm.full[s], kb * BK, tc.n0, tc.g);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            constexpr uint32_t idesc = make_instr_desc<BN>();
            uint32_t it = 0, local = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
                const uint32_t acc = local & 1;
                mbar_wait(&sm.acc_empty[acc], ((local >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&sm.full[s], (it / STAGES) & 1);
                    tcgen05_fence_after();
                    const uint64_t da = make_smem_desc(smem_u32(sm.a[s])), db = make_smem_desc(smem_u32(sm.b[s]));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 bytes (16 bf16) along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
                        umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tcgen05_commit(&sm.empty[s]);  // arrives when the MMAs above have finished reading this stage
                }
                tcgen05_commit(&sm.acc_full[acc]);
            }
        }
    } else {  // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t local = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
            const TileCoord tc = tile_coord(tile, tiles_m, tiles_n, BN);
            const uint32_t acc = local & 1;
            mbar_wait(&sm.acc_full[acc], (local >> 1) & 1);
            tcgen05_fence_after();
            __nv_bfloat16* crow;
            if constexpr (PEER) {
                const int src = tc.m0 / peer.cap, c0 = tc.m0 - src * peer.cap;
                crow = reinterpret_cast<__nv_bfloat16*>(peer.ptr[src] + peer.off) +
                       ((static_cast<size_t>(peer.rank) * peer.groups + tc.g) * peer.cap + (c0 + row)) * N + tc.n0;
            } else {
                crow = C + (static_cast<size_t>(tc.g) * M + (tc.m0 + row)) * N + tc.n0;
            }
            const float* brow = bias ? bias + static_cast<size_t>(tc.g) * N + tc.n0 : nullptr;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + static_cast<uint32_t>(c), r);
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v0 = __uint_as_float(r[2 * j]), v1 = __uint_as_float(r[2 * j + 1]);
                    if (brow) v0 += __ldg(brow + c + 2 * j), v1 += __ldg(brow + c + 2 * j + 1);
                    if (act == 1) v0 = gelu_tanh(v0), v1 = gelu_tanh(v1);
                    __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                    packed[j] = *reinterpret_cast<uint32_t*>(&h);
                }
                uint4* dst = reinterpret_cast<uint4*>(crow + c);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.acc_empty[acc]);  // this warp no longer reads the accumulator
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
    }
}

// =====================================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile. CTA r holds rows
// [m0 + 128 r, +128) of A and rows [n0 + 128 r, +128) of B in its shared memory; ONE thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M256 N256 K16) which reads A from "its own" SM and both halves of B from both SMs, so each
// SM stages only half of B per K block (32 KB per stage instead of 48 KB → 6 stages) and the tensor pipes of both SMs are
// driven by a single instruction stream. Accumulators: 128 lanes x 256 columns in each CTA's TMEM, double buffered.
//   full[s]      (leader's)  : 2 arrivals (leader's arrive.expect_tx + peer's remote arrive) + bytes of all four TMA loads
//   empty[s]     (both CTAs) : multicast tcgen05.commit from the leader's MMA thread
//   acc_full[a]  (both CTAs) : multicast tcgen05.commit
//   acc_empty[a] (leader's)  : 8 arrivals, one per epilogue warp of the pair (the peer's arrive remotely)
// Opt-in (BAGUA_GEMM_2CTA=1): written without hardware access, numerics test gated by BAGUA_EXPERIMENTAL=1.
// =====================================================================================================================
constexpr int BN2 = 256;
constexpr int kStages2 = 6;
constexpr uint32_t kStageBytesB2 = (BN2 / 2) * BK * 2;

struct __align__(1024) GemmSmem2 {
    uint8_t a[kStages2][kStageBytesA];
    uint8_t b[kStages2][kStageBytesB2];
    uint64_t full[kStages2];
    uint64_t empty[kStages2];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(p)), "r"(rank));
    return out;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 0b11;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
    grouped_gemm_tn_2cta_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, __nv_bfloat16* __restrict__ C,
                                const float* __restrict__ bias, int M, int N, int K, int num_tiles, int tiles_m, int tiles_n, int act) {
    extern __shared__ uint8_t smem_raw[];
    GemmSmem2& sm = *reinterpret_cast<GemmSmem2*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = K / BK;
    const uint32_t cta = cluster_ctarank();
    const bool leader = cta == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    constexpr int TM = 2 * BM;  // rows per cluster tile

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
        for (int s = 0; s < kStages2; ++s) {
            mbar_init(&sm.full[s], 2);
            mbar_init(&sm.empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.acc_full[i], 1);
            mbar_init(&sm.acc_empty[i], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the same warp of both CTAs allocates (and later frees) the pair's TMEM columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(2 * BN2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();  // barrier inits and TMEM base visible to both CTAs
    tcgen05_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs: own half of A, own half of B; bytes are credited to the leader's barrier) =====
            uint32_t it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = (tile % tiles_m) * TM;
                const int r = tile / tiles_m;
                const int n0 = (r % tiles_n) * BN2, g = r / tiles_n;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kStages2;
                    mbar_wait(&sm.empty[s], ((it / kStages2) & 1) ^ 1);
                    const uint32_t lead_full = mapa_u32(&sm.full[s], 0);
                    if (leader)
                        mbar_expect_tx(&sm.full[s], 2 * (kStageBytesA + kStageBytesB2));
                    else
                        mbar_arrive_cluster(lead_full);
                    tma_load_3d_2sm(sm.a[s], &tm_a, lead_full, kb * BK, m0 + static_cast<int>(cta) * BM, g);
                    tma_load_3d_2sm(sm.b[s], &tm_b, lead_full, kb * BK, n0 + static_cast<int>(cta) * (BN2 / 2), g);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {  // ===== MMA issuer: one thread for both SMs =====
            // instruction descriptor as in the 1-CTA kernel with M = 256
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN2 >> 3) << 17) | (static_cast<uint32_t>(TM >> 4) << 24);
            uint32_t it = 0, local = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
                const uint32_t acc = local & 1;
                mbar_wait(&sm.acc_empty[acc], ((local >> 1) & 1) ^ 1);  // all 8 epilogue warps of the pair have drained it
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN2;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kStages2;
                    mbar_wait(&sm.full[s], (it / kStages2) & 1);
                    tcgen05_fence_after();
                    const uint64_t da = make_smem_desc(smem_u32(sm.a[s])), db = make_smem_desc(smem_u32(sm.b[s]));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16_2sm(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    tcgen05_commit_2sm(&sm.empty[s]);  // frees stage s in BOTH CTAs
                }
                tcgen05_commit_2sm(&sm.acc_full[acc]);  // wakes the epilogue warps of BOTH CTAs
            }
        }
    } else {  // ===== epilogue warps 2..5 of each CTA: its own 128 rows =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t local = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
            const int m0 = (tile % tiles_m) * TM + static_cast<int>(cta) * BM;
            const int r = tile / tiles_m;
            const int n0 = (r % tiles_n) * BN2, g = r / tiles_n;
            const uint32_t acc = local & 1;
            mbar_wait(&sm.acc_full[acc], (local >> 1) & 1);
            tcgen05_fence_after();
            __nv_bfloat16* crow = C + (static_cast<size_t>(g) * M + (m0 + row)) * N + n0;
            const float* brow = bias ? bias + static_cast<size_t>(g) * N + n0 : nullptr;
#pragma unroll 1
            for (int c = 0; c < BN2; c += 32) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN2 + static_cast<uint32_t>(c), rr);
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v0 = __uint_as_float(rr[2 * j]), v1 = __uint_as_float(rr[2 * j + 1]);
                    if (brow) v0 += __ldg(brow + c + 2 * j), v1 += __ldg(brow + c + 2 * j + 1);
                    if (act == 1) v0 = gelu_tanh(v0), v1 = gelu_tanh(v1);
                    __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                    packed[j] = *reinterpret_cast<uint32_t*>(&h);
                }
                uint4* dst = reinterpret_cast<uint4*>(crow + c);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&sm.acc_empty[acc], 0));  // the MMA thread waits on the LEADER's barrier
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();  // neither CTA may free TMEM / exit while its partner still reads operands or accumulators
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN2) : "memory");
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) throw std::runtime_error("bagua: cuTensorMapEncodeTiled is not available from the driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [G, rows, K] bf16, K contiguous → box (BK x box_rows x 1), 128-byte swizzle
CUtensorMap make_map(const void* ptr, int G, int rows, int K, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(G)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(K) * 2, static_cast<cuuint64_t>(rows) * K * 2};
    cuuint32_t box[3] = {BK, static_cast<cuuint32_t>(box_rows), 1};  // ≤ 256 rows per box
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("bagua: cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
    return m;
}

}  // namespace

bool grouped_gemm_supported(int M, int N, int K) { return M > 0 && N > 0 && K > 0 && M % BM == 0 && N % 128 == 0 && K % BK == 0; }

namespace {
template <int BN, bool PEER>
void launch_bn(const void* A, const void* B, void* C, const float* bias, int G, int M, int N, int K, int act, const GemmPeerOut& peer,
               cudaStream_t stream) {
    const CUtensorMap ta = make_map(A, G, M, K, BM), tb = make_map(B, G, N, K, BN);
    const size_t smem = sizeof(GemmSmem<BN>) + 1024;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        BAGUA_CUDA_CHECK(cudaFuncSetAttribute(grouped_gemm_tn_kernel<BN, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int dev = 0;
        BAGUA_CUDA_CHECK(cudaGetDevice(&dev));
        BAGUA_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const int tiles_m = M / BM, tiles_n = N / BN;
    const int num_tiles = tiles_m * tiles_n * G;
    const int grid = num_tiles < num_sms ? num_tiles : num_sms;  // persistent: one CTA per SM
    grouped_gemm_tn_kernel<BN, PEER><<<grid, kGemmThreads, smem, stream>>>(ta, tb, static_cast<__nv_bfloat16*>(C), bias, M, N, K, num_tiles, tiles_m,
                                                                            tiles_n, act, peer);
}
}  // namespace

namespace {
bool use_2cta(int M, int N) {
    static const bool enabled = [] {
        const char* v = getenv("BAGUA_GEMM_2CTA");
        return v && v[0] == '1';
    }();
    return enabled && M % 256 == 0 && N % 256 == 0;
}

void launch_2cta(const void* A, const void* B, void* C, const float* bias, int G, int M, int N, int K, int act, cudaStream_t stream) {
    const CUtensorMap ta = make_map(A, G, M, K, BM), tb = make_map(B, G, N, K, BN2 / 2);
    const size_t smem = sizeof(GemmSmem2) + 1024;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        BAGUA_CUDA_CHECK(cudaFuncSetAttribute(grouped_gemm_tn_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int dev = 0;
        BAGUA_CUDA_CHECK(cudaGetDevice(&dev));
        BAGUA_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const int tiles_m = M / (2 * BM), tiles_n = N / BN2;
    const int num_tiles = tiles_m * tiles_n * G;
    const int clusters = num_tiles < num_sms / 2 ? num_tiles : num_sms / 2;  // persistent: one CTA pair per TPC
    grouped_gemm_tn_2cta_kernel<<<2 * clusters, kGemmThreads, smem, stream>>>(ta, tb, static_cast<__nv_bfloat16*>(C), bias, M, N, K, num_tiles, tiles_m,
                                                                              tiles_n, act);
}
}  // namespace

void launch_grouped_gemm_tn(const void* A, const void* B, void* C, const float* bias, int G, int M, int N, int K, int act, cudaStream_t stream) {
    if (!grouped_gemm_supported(M, N, K)) throw std::runtime_error("bagua: grouped_gemm_tn needs M%128==0, N%128==0, K%64==0");
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15u)
        throw std::runtime_error("bagua: grouped_gemm_tn needs 16-byte aligned operands");
    const GemmPeerOut none{};
    if (use_2cta(M, N))
        launch_2cta(A, B, C, bias, G, M, N, K, act, stream);
    else if (N % 256 == 0)
        launch_bn<256, false>(A, B, C, bias, G, M, N, K, act, none, stream);
    else
        launch_bn<128, false>(A, B, C, bias, G, M, N, K, act, none, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("bagua: launch of grouped_gemm_tn failed: ") + cudaGetErrorString(e));
    count_launch();
}

// GEMM whose epilogue IS the MoE combine all-to-all: A [G, world*cap, K] (rows ordered [source rank][slot]), B [G, N, K];
// the 128-row tiles of source rank s are stored into rank s's symmetric buffer at [this rank][g][slot][N].
// No cross-GPU synchronisation inside: the consumer kernel (moe_gather, local layout) starts with the peer barrier that
// makes the pushes of every owner visible, and ends with the one that lets the owners reuse the buffer.
void launch_grouped_gemm_tn_push(const void* A, const void* B, const float* bias, int G, int N, int K, int act, const PeerCtx& ctx, const PeerBuf& out,
                                 size_t out_off, int cap, cudaStream_t stream) {
    const int M = ctx.world * cap;
    if (!grouped_gemm_supported(M, N, K) || cap % BM != 0) throw std::runtime_error("bagua: grouped_gemm_tn_push needs cap%128==0, N%128==0, K%64==0");
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | out_off) & 15u)
        throw std::runtime_error("bagua: grouped_gemm_tn_push needs 16-byte aligned operands");
    GemmPeerOut peer{};
    for (int p = 0; p < ctx.world; ++p) peer.ptr[p] = out.ptr[p];
    peer.off = out_off;
    peer.rank = ctx.rank;
    peer.cap = cap;
    peer.groups = G;
    if (N % 256 == 0)
        launch_bn<256, true>(A, B, nullptr, bias, G, M, N, K, act, peer, stream);
    else
        launch_bn<128, true>(A, B, nullptr, bias, G, M, N, K, act, peer, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("bagua: launch of grouped_gemm_tn_push failed: ") + cudaGetErrorString(e));
    count_launch();
}

}  // namespace bagua
