// evg_policy_mlp.cu — the policy-in-the-loop forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// The only contraction on any configuration's path (BASELINE.json configs[3]): DQN's
//     Q = Linear(105, 528) - ReLU - Linear(528, 132)          agents/DQN/QNetwork.py:37,42
// evaluated for every (match, player) row of the observation tensor the step kernel has just written.  One kernel does
// both layers; the hidden activations never leave the SM:
//
//   CTA = 512 threads = one tile of 128 observation rows at a time (UMMA M = 128: TMEM lane i holds row i of the
//   accumulators; four threads per row, one per quarter of the columns), persistent over the tiles.
//   per tile:  A  <- the tile's float32 observations, converted to bf16 into the K-major 128-byte-swizzled layout
//              for each chunk c of 192 hidden units (528 -> 3 chunks, zero padded):
//                 D1[128 x 192]  = A[128 x 128] . W1c^T          8 x tcgen05.mma  (K = 16 each), accumulators in TMEM
//                 H              = bf16(relu(D1))                tcgen05.ld -> registers -> swizzled shared memory
//                 D2[128 x 144] += H[128 x 192] . W2c^T          12 x tcgen05.mma, accumulating over the chunks in TMEM
//              Q <- D2                                             tcgen05.ld -> registers -> global (float32)
//   Both biases ride inside the GEMMs (homogeneous coordinates): A carries a column of ones behind the obs_len features
//   whose weights are b1, and one padding hidden unit is wired to be relu(1) = 1 with b2 as its outgoing weights
//   (evgsim.policy.pack_mlp builds the images that way), so the epilogues are a ReLU and a store.
//   The weight chunks W1c / W2c are bf16 IMAGES of the shared-memory operand layout, fetched by the bulk-copy engine
//   (cp.async.bulk, one instruction per 48 / 54 KB image, completion on an mbarrier) as soon as the MMAs that read the
//   buffer's previous content have retired: W1(c+1) arrives under the activation epilogue and layer 2 of chunk c,
//   W2(c+1) under layer 1 and the epilogue of chunk c+1.  One elected thread issues copies and MMAs and commits the
//   MMAs to mbarriers (tcgen05.commit); the CTA waits on those.
//
// bf16 operands, fp32 accumulation.  Q is written either row-major [rows][out] (torch's layout) or transposed
// [out][rows] — what the thread-per-row decode (evg_decode_dqn, pinned to the reference's DQNAgent.filter_actions)
// reads coalesced.  tests/test_gpu_policy.py checks Q against torch fp32 within the bf16 tolerance stated there.
// Weight images are built on the host by evgsim.policy.pack_mlp (same swizzle function as below).
#include <cuda_bf16.h>

#include "evg_internal.h"

namespace evg {

namespace {

constexpr int kMlpThreads = 512;
constexpr int kTileM = 128;   // observation rows per tile
constexpr int kInPad = 128;   // input features, padded (obs_len <= 128)
constexpr int kChunk = 192;   // hidden units per chunk: 3 swizzle atoms of 64
constexpr int kOutPad = 144;  // outputs, padded to a multiple of 16 (12 groups x 11 nodes = 132)
constexpr int kAtomK = 64;    // bf16 elements in one 128-byte swizzle row

// shared-memory carve-up (bytes; every operand block starts on a 1024-byte boundary, as the 128-byte swizzle needs)
constexpr int kSmA = 0;                                            // 2 atoms x [128 rows x 128 B]
constexpr int kSmH = kSmA + (kInPad / kAtomK) * kTileM * 128;      // 3 atoms x [128 rows x 128 B]
constexpr int kSmW1 = kSmH + (kChunk / kAtomK) * kTileM * 128;     // 2 atoms x [192 rows x 128 B]
constexpr int kSmW2 = kSmW1 + (kInPad / kAtomK) * kChunk * 128;    // 3 atoms x [144 rows x 128 B]
constexpr int kSmBar = kSmW2 + (kChunk / kAtomK) * kOutPad * 128;  // 4 mbarriers + the TMEM base address
constexpr int kSmBytes = kSmBar + 48;
constexpr int kW1ChunkBytes = (kInPad / kAtomK) * kChunk * 128;    // 49152
constexpr int kW2ChunkBytes = (kChunk / kAtomK) * kOutPad * 128;   // 55296
static_assert(kSmH % 1024 == 0 && kSmW1 % 1024 == 0 && kSmW2 % 1024 == 0 && (kOutPad * 128) % 1024 == 0 && (kChunk * 128) % 1024 == 0, "swizzle atoms must be 1024-byte aligned");

constexpr int kTmemCols = 512;  // power of two >= D1 (192) + D2 (144) columns
constexpr int kTmemD1 = 0, kTmemD2 = 256;

// K-major operand tile with 128-byte swizzle: element (row r, column k) of a [rows x K] bf16 matrix lives in atom k / 64
// (a block of rows x 128 bytes), at row r, 16-byte chunk ((k % 64) / 8) ^ (r % 8)
__host__ __device__ inline int swz_offset(int rows, int r, int k)
{
    return (k / kAtomK) * rows * 128 + r * 128 + ((((k % kAtomK) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 [0:14), leading byte offset >> 4
// [16:30) (unused for a swizzled K-major operand), stride byte offset >> 4 [32:46) = 1024 B between 8-row groups,
// version 1 [46:48), layout SWIZZLE_128B = 2 [61:64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (uint64_t)1 << 16 | (uint64_t)(1024 >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)2 << 61;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4:6) = 1, A bf16 [7:10) = 1, B bf16 [10:13) = 1, both
// K-major, N >> 3 at [17:23), M >> 4 at [24:29)
__device__ __forceinline__ uint32_t make_idesc(int m, int n)
{
    return 1u << 4 | 1u << 7 | 1u << 10 | (uint32_t)(n >> 3) << 17 | (uint32_t)(m >> 4) << 24;
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// wait for phase `parity` of an mbarrier; a bounded spin that traps instead of hanging the GPU if the MMAs never arrive
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 26)) __trap();
    }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}

// one bulk copy global -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

// q_col_stride == 1: Q[row * q_row_stride + col] (row-major); q_row_stride == 1: Q[col * q_col_stride + row] (transposed)
__global__ void __launch_bounds__(kMlpThreads, 1)
evg_policy_mlp_kernel(const float* __restrict__ obs, int64_t rows, int in_dim, const unsigned char* __restrict__ w1_img,
                      const unsigned char* __restrict__ w2_img, int n_chunks, int out_dim, float* __restrict__ q, int64_t q_row_stride,
                      int64_t q_col_stride)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSmBar);  // [0] layer-1 MMAs done, [1] layer-2 MMAs done, [2] W1 image landed, [3] W2 image landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmBar + 32);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // one warp allocates the tensor memory (and frees it at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // my row = TMEM lane 32 * (warp % 4) + lane (a warp reaches the lane quarter warp % 4); my quarter of the columns = warp / 4
    const int r = (warp & 3) * 32 + lane, quarter = warp >> 2;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t idesc1 = make_idesc(kTileM, kChunk), idesc2 = make_idesc(kTileM, kOutPad);
    const uint32_t sA = smem_u32(smem + kSmA), sH = smem_u32(smem + kSmH), sW1 = smem_u32(smem + kSmW1), sW2 = smem_u32(smem + kSmW2);
    const int64_t n_tiles = (rows + kTileM - 1) / kTileM;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_steps = my_tiles * n_chunks;  // (tile, chunk) steps of this CTA: the weight images cycle through the chunks
    uint32_t ph_m1 = 0, ph_m2 = 0, ph_w1 = 0, ph_w2 = 0;
    if (tid == 0 && n_steps > 0) {
        bulk_load(smem + kSmW1, w1_img, kW1ChunkBytes, &bar[2]);
        bulk_load(smem + kSmW2, w2_img, kW2ChunkBytes, &bar[3]);
    }
    // A: a tile's observations -> bf16, swizzled.  Every thread owns 4 of the tile's 2048 16-byte chunks (8 features each):
    // all 32 loads are issued together into registers (one DRAM round trip), converted and stored later; feature
    // `in_dim` is the constant 1 that carries the bias, rows and features beyond that are zero.
    float af[kTileM * (kInPad / 8) / kMlpThreads][8];
    auto load_a = [&](int64_t row0, int nrows) {
#pragma unroll
        for (int it = 0; it < kTileM * (kInPad / 8) / kMlpThreads; ++it) {
            const int i = tid + it * kMlpThreads, ar = i >> 4, k0 = (i & 15) * 8;
            const float* src = obs + (row0 + ar) * in_dim + k0;
#pragma unroll
            for (int e = 0; e < 8; ++e) af[it][e] = (ar < nrows && k0 + e < in_dim) ? __ldcs(src + e) : (k0 + e == in_dim ? 1.f : 0.f);
        }
    };
    auto store_a = [&]() {
#pragma unroll
        for (int it = 0; it < kTileM * (kInPad / 8) / kMlpThreads; ++it) {
            const int i = tid + it * kMlpThreads, ar = i >> 4, k0 = (i & 15) * 8;
            *reinterpret_cast<uint4*>(smem + kSmA + swz_offset(kTileM, ar, k0)) =
                make_uint4(pack_bf16(af[it][0], af[it][1]), pack_bf16(af[it][2], af[it][3]), pack_bf16(af[it][4], af[it][5]), pack_bf16(af[it][6], af[it][7]));
        }
    };
    int64_t step = 0;
    bool l1_queued = false;  // the coming tile's first layer-1 MMAs were issued at the end of the previous tile
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTileM;
        const int nrows = rows - row0 < kTileM ? (int)(rows - row0) : kTileM;
        if (tile == (int64_t)blockIdx.x) {  // the first tile: later ones are fetched and converted under the previous tile's last chunk
            load_a(row0, nrows);
            store_a();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core's reads
            __syncthreads();
        }
        for (int c = 0; c < n_chunks; ++c, ++step) {
            const int cn = c + 1 == n_chunks ? 0 : c + 1;  // the chunk of the next step
            // ---- layer 1: D1 = A . W1c^T, as soon as W1c has landed.  Only a tile's first chunk is issued here: the later
            // ones are queued right behind the previous chunk's layer 2 (below), so that their latency is not paid again
            auto issue_layer1 = [&](uint32_t parity) {
                mbar_wait(&bar[2], parity);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < kInPad / 16; ++k) {  // within a swizzle atom the start address advances by 32 bytes per K = 16
                    const uint32_t offA = (k >> 2) * (kTileM * 128) + (k & 3) * 32, offB = (k >> 2) * (kChunk * 128) + (k & 3) * 32;
                    umma(tmem + kTmemD1, make_desc(sA + offA), make_desc(sW1 + offB), idesc1, k > 0);
                }
                umma_commit(&bar[0]);
            };
            if (tid == 0 && c == 0 && !l1_queued) issue_layer1(ph_w1);
            l1_queued = false;
            ph_w1 ^= 1;
            mbar_wait(&bar[0], ph_m1);  // D1 complete (and, the tensor pipe being in order, the previous step's layer 2: H is free)
            ph_m1 ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0 && step + 1 < n_steps) bulk_load(smem + kSmW1, w1_img + (size_t)cn * kW1ChunkBytes, kW1ChunkBytes, &bar[2]);
            const bool fetch_next_a = c + 1 == n_chunks && tile + gridDim.x < n_tiles;  // this tile's last layer-1 MMAs have read A: it is free
            if (fetch_next_a) {
                const int64_t nrow0 = (tile + gridDim.x) * kTileM;
                load_a(nrow0, rows - nrow0 < kTileM ? (int)(rows - nrow0) : kTileM);  // in flight under the epilogue below
            }
            // ---- hidden activations of my row, my 48 of the chunk's 192 columns: TMEM -> registers -> ReLU, bf16 -> swizzled H
            {
                constexpr int kLd = kChunk / 4 / 16;  // tensor-memory loads of 16 columns per thread: all issued, then one wait
                uint32_t v[kLd][16];
                const int jq1 = quarter * (kChunk / 4);
#pragma unroll
                for (int t = 0; t < kLd; ++t) tmem_ld16(lane_base + kTmemD1 + jq1 + 16 * t, v[t]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int t = 0; t < kLd; ++t) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {  // 8 hidden units = one 16-byte chunk of the swizzled row
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            w[e] = pack_bf16(fmaxf(__uint_as_float(v[t][8 * g + 2 * e]), 0.f), fmaxf(__uint_as_float(v[t][8 * g + 2 * e + 1]), 0.f));
                        *reinterpret_cast<uint4*>(smem + kSmH + swz_offset(kTileM, r, jq1 + 16 * t + 8 * g)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            // ---- layer 2: D2 += H . W2c^T
            if (tid == 0) {
                mbar_wait(&bar[3], ph_w2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < kChunk / 16; ++k) {
                    const uint32_t offA = (k >> 2) * (kTileM * 128) + (k & 3) * 32, offB = (k >> 2) * (kOutPad * 128) + (k & 3) * 32;
                    umma(tmem + kTmemD2, make_desc(sH + offA), make_desc(sW2 + offB), idesc2, (c > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&bar[1]);
                // the next chunk's layer 1 behind it: every thread has finished reading D1 (barrier above), the tensor pipe
                // runs in order, and W1's next chunk was requested when this chunk's layer 1 completed
                if (c + 1 < n_chunks) issue_layer1(ph_w1);
            }
            ph_w2 ^= 1;
            if (fetch_next_a) {  // the next tile's A, converted while layer 2 runs
                store_a();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                // ... and its first layer-1 MMAs queued now: they only write D1, so they run under this tile's Q store
                if (tid == 0) issue_layer1(ph_w1);
                l1_queued = true;
            }
            mbar_wait(&bar[1], ph_m2);  // the W2 buffer is free again; after the last chunk D2 is complete
            ph_m2 ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0 && step + 1 < n_steps) bulk_load(smem + kSmW2, w2_img + (size_t)cn * kW2ChunkBytes, kW2ChunkBytes, &bar[3]);
        }
        // ---- Q = D2: my row, my quarter of the columns (36 = 16 + 16 + 4)
        {
            float* qp = q + (row0 + r) * q_row_stride + (int64_t)(quarter * (kOutPad / 4)) * q_col_stride;
            const int jq = quarter * (kOutPad / 4);
            const bool live = r < nrows;
            uint32_t v[2][16], v4[4];
            tmem_ld16(lane_base + kTmemD2 + jq, v[0]);
            tmem_ld16(lane_base + kTmemD2 + jq + 16, v[1]);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v4[0]), "=r"(v4[1]), "=r"(v4[2]), "=r"(v4[3]) : "r"(lane_base + kTmemD2 + jq + 32));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 16) {
#pragma unroll
                for (int e = 0; e < 16; ++e, qp += q_col_stride)
                    if (live && jq + j0 + e < out_dim) *qp = __uint_as_float(v[j0 >> 4][e]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e, qp += q_col_stride)
                if (live && jq + 32 + e < out_dim) *qp = __uint_as_float(v4[e]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // the next tile's first MMAs overwrite D1 / D2, its conversion overwrites A
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

}  // namespace

cudaError_t launch_policy_mlp(const float* obs, int64_t rows, int in_dim, const void* w1_img, const void* w2_img, int n_chunks, int out_dim, float* q,
                              int transposed, int sm_count, cudaStream_t stream)
{
    if (rows <= 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(evg_policy_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int64_t tiles = (rows + kTileM - 1) / kTileM;
    const unsigned grid = (unsigned)(tiles < sm_count ? tiles : sm_count);
    evg_policy_mlp_kernel<<<grid, kMlpThreads, kSmBytes, stream>>>(obs, rows, in_dim, reinterpret_cast<const unsigned char*>(w1_img),
                                                                   reinterpret_cast<const unsigned char*>(w2_img), n_chunks, out_dim, q,
                                                                   transposed ? 1 : out_dim, transposed ? rows : 1);
    return cudaGetLastError();
}

}  // namespace evg
