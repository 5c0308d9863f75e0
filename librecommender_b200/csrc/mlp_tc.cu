// Dense layer on the 5th-generation tensor cores with fp32-level accuracy: Y = act(X Wt^T + b).
//
// Replaces tf_dense (reference libreco/layers/dense.py:52-80, BN folded by the caller) for the MLP
// tails of DeepFM / DIN / YouTubeRanking / TwoTower when the layer is large enough to be
// compute-bound on the SIMT path (b200_linear_f32).  The reference computes these layers in fp32
// (TensorFlow MatMul); a plain tf32 or bf16 tensor-core GEMM would miss the 1e-5 parity bar, so
// the operands are split  x = hi + lo  (hi = top 11 mantissa bits, lo = next 11) and three
// tcgen05.mma kind::tf32 products are accumulated:  hi*hi + lo*hi + hi*lo  (truncation keeps 10
// explicit mantissa bits, |lo| < 2^-10 |x|: the dropped lo*lo term is < 2^-20 |x||w| and the two
// truncated cross terms add < 2^-20 each; tests/test_tf32x3_model_cpu.py).  The tensor core truncates its fp32 accumulator once per MMA (measured: a
// single accumulator over 48 MMAs drifts by ~2e-6 of sum|x w|), so (a) the dominant hi*hi products
// and the 2^-11-times smaller cross products go to SEPARATE TMEM accumulators, and (b) both are
// promoted to registers (round-to-nearest adds) every GC k-chunks = GC*4 MMAs per accumulator.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: X tile [128 x 32 fp32] and Wt tile [n_pad x 32 fp32] per k-chunk
//   warp 1      MMA issuer (one elected lane)
//   warps 2-5   splitters: write the `lo` tile next to every landed X tile (the X tile itself is the
//               `hi` operand); the weight tiles arrive pre-split when the caller made a split copy
//   warps 6-9   epilogue: TMEM -> registers (+=), then bias / ReLU / store at the end of a row tile
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/b200reco.h"

namespace b200 {
namespace mlp {

constexpr int TM = 128;          // rows per tile (UMMA M)
constexpr int KC = 32;           // fp32 per k-chunk = one 128-byte swizzled row
constexpr int GC = 2;            // k-chunks per accumulator group (promotion interval = 64 k)
constexpr int NMAX = 128;        // output columns per CTA (grid.y covers wider layers)
constexpr int MAXSTAGE = 4;
constexpr int THREADS = 320;
constexpr int A_BYTES = TM * KC * 4;   // 16 KB

struct Params {
  int64_t R;
  int64_t ldy;
  const float* bias;
  float* Y;
  const float* dot_w;     // fused DIN epilogue: 16 weights of the Dense(1) on sigmoid(Dense(16)) (NULL = plain layer)
  int din, dout, n_pad, relu;
  int n_tiles, n_chunks, nstage;
  int n_chunks_total;      // split-K: CTA z takes k-chunks [z * n_chunks, min((z + 1) * n_chunks, n_chunks_total))
  int64_t split_stride;    //          and writes its partial product to Y + z * split_stride (no bias / ReLU)
};

struct Smem {
  uint64_t full[MAXSTAGE], split[MAXSTAGE], empty[MAXSTAGE];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad_[3];
  float bias_s[NMAX];      // bias of this CTA's column block (0 past dout / without bias): the epilogue read 128
                           // separate global words per thread and tile before — 27 % of the kernel's stall samples
};

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  // D = f32 (1 @ bit 4), A = B = tf32 (2 @ bits 7, 10), both K-major, dense
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// x -> (hi, lo): hi keeps the 10 explicit mantissa bits of tf32, lo = x - hi (exact in fp32).
// The tensor core reads 32-bit containers and ignores the 13 low mantissa bits (truncation —
// verified by tests/test_gpu_linear_tc.py: a rounding conversion would show up as a 2^-11 error),
// so the X tile itself serves as the `hi` operand and only the `lo` tile is written.
__device__ __forceinline__ float lo_part(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}
__device__ __forceinline__ float4 lo4(const float4 v) {
  return make_float4(lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w));
}

// weights: explicit hi / lo copies, made once per layer (b200_linear_tf32x3_split_weights)
__global__ void split_weights_kernel(const float* __restrict__ W, int64_t ldw, int din, int dout, int64_t ld,
                                     float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)dout * ld) return;
  const int r = (int)(i / ld), c = (int)(i % ld);
  const float x = c < din ? W[(int64_t)r * ldw + c] : 0.f;
  const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  out[i] = h;
  out[(int64_t)dout * ld + i] = x - h;
}

// WSPLIT: tmW / tmWlo address the pre-split weight copies; otherwise the splitters also split the
// weight tile of every stage (self-contained call, more shared-memory traffic).
template <bool WSPLIT, bool DOT>
__global__ void __launch_bounds__(THREADS, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmWlo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.n_pad * KC * 4;
  const int stage_bytes = 2 * A_BYTES + 2 * b_bytes;
  Smem* ss = (Smem*)(smem + (size_t)p.nstage * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int col0 = blockIdx.y * NMAX;
  const int kc_base = blockIdx.z * p.n_chunks;
  const int kc_count = min(p.n_chunks, p.n_chunks_total - kc_base);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstage; ++s) {
      ptx::mbar_init(&ss->full[s], 1);
      ptx::mbar_init(&ss->split[s], 4);
      ptx::mbar_init(&ss->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&ss->tmem_full[a], 1);
      ptx::mbar_init(&ss->tmem_empty[a], 4);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
  }
  if (warp == 1) {
    ptx::tmem_alloc(&ss->tmem_base, 4 * NMAX);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + NMAX) {
    const int c = threadIdx.x - 64;
    ss->bias_s[c] = (p.bias && col0 + c < p.dout) ? __ldg(p.bias + col0 + c) : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ss->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < kc_count; ++kc) {
          ptx::mbar_wait(&ss->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&ss->full[stage], (uint32_t)(A_BYTES + (WSPLIT ? 2 : 1) * b_bytes));
          uint8_t* st = smem + (size_t)stage * stage_bytes;
          const int kx = (kc_base + kc) * KC;
          ptx::tma_load_2d(st, &tmX, &ss->full[stage], kx, tile * TM);
          ptx::tma_load_2d(st + 2 * A_BYTES, &tmW, &ss->full[stage], kx, col0);
          if (WSPLIT) ptx::tma_load_2d(st + 2 * A_BYTES + b_bytes, &tmWlo, &ss->full[stage], kx, col0);
          if (++stage == p.nstage) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(TM, p.n_pad);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t s_addr = ptx::smem_u32(smem);
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < kc_count; ++kc) {
          const int gpos = kc % GC;
          if (gpos == 0) {   // new accumulator group
            ptx::mbar_wait(&ss->tmem_empty[acc], acc_phase ^ 1);
            ptx::tc_fence_after();
          }
          ptx::mbar_wait(&ss->split[stage], phase);
          ptx::tc_fence_after();
          const uint32_t st = s_addr + (uint32_t)(stage * stage_bytes);
          const uint64_t a_hi = ptx::umma_desc_sw128_kmajor(st);
          const uint64_t a_lo = ptx::umma_desc_sw128_kmajor(st + A_BYTES);
          const uint64_t b_hi = ptx::umma_desc_sw128_kmajor(st + 2 * A_BYTES);
          const uint64_t b_lo = ptx::umma_desc_sw128_kmajor(st + 2 * A_BYTES + b_bytes);
          const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * NMAX);   // hi*hi
          const uint32_t d_corr = d_main + NMAX;                            // lo*hi + hi*lo
#pragma unroll
          for (int k4 = 0; k4 < KC / 8; ++k4) {
            // 8 tf32 = 32 bytes per MMA inside the 128-byte swizzled row: +2 in the >>4 field
            const uint64_t o = (uint64_t)(k4 * 2);
            const uint32_t cont = (uint32_t)((gpos | k4) != 0);
            umma_tf32(d_main, a_hi + o, b_hi + o, idesc, cont);
            umma_tf32(d_corr, a_lo + o, b_hi + o, idesc, cont);
            umma_tf32(d_corr, a_hi + o, b_lo + o, idesc, 1u);
          }
          ptx::umma_commit(&ss->empty[stage]);
          if (++stage == p.nstage) { stage = 0; phase ^= 1; }
          if (gpos == GC - 1 || kc == kc_count - 1) {
            ptx::umma_commit(&ss->tmem_full[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== splitters =====================
    const int t = threadIdx.x - 64;          // 0..127
    int stage = 0;
    uint32_t phase = 0;
    const int b_vec = b_bytes / 16;          // float4 per B tile
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < kc_count; ++kc) {
        ptx::mbar_wait(&ss->full[stage], phase);
        uint8_t* st = smem + (size_t)stage * stage_bytes;
        float4* ah = (float4*)st;
        float4* al = (float4*)(st + A_BYTES);
#pragma unroll
        for (int j = 0; j < A_BYTES / 16 / 128; ++j) al[t + j * 128] = lo4(ah[t + j * 128]);
        if (!WSPLIT) {
          const float4* bh = (const float4*)(st + 2 * A_BYTES);
          float4* bl = (float4*)(st + 2 * A_BYTES + b_bytes);
          for (int j = t; j < b_vec; j += 128) bl[j] = lo4(bh[j]);
        }
        ptx::fence_proxy_async_smem();       // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ss->split[stage]);
        if (++stage == p.nstage) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                  // TMEM lane quadrant this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    const int n_groups = (kc_count + GC - 1) / GC;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      float y[NMAX];
#pragma unroll
      for (int i = 0; i < NMAX; ++i) y[i] = 0.f;
      for (int g = 0; g < n_groups; ++g) {
        ptx::mbar_wait(&ss->tmem_full[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * NMAX);
#pragma unroll
        for (int c = 0; c < NMAX / 32; ++c) {
          if (c * 32 < p.n_pad) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(taddr + (uint32_t)(NMAX + c * 32), r);     // correction first
            ptx::tmem_ld_wait_regs(r);
#pragma unroll
            for (int i = 0; i < 32; ++i) y[c * 32 + i] += __uint_as_float(r[i]);
            ptx::tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), r);
            ptx::tmem_ld_wait_regs(r);
#pragma unroll
            for (int i = 0; i < 32; ++i) y[c * 32 + i] += __uint_as_float(r[i]);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ss->tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      const int64_t row = (int64_t)tile * TM + q * 32 + lane;
      if (DOT && row < p.R) {
        // fused attention epilogue (DIN all-items): per group of 16 columns ONE output
        //   a[row, (col0 + 16 g) / 16] = sum_j dot_w[j] * sigmoid(y[16 g + j] + bias)
        // the [R, dout] pre-activations (16x the bytes) are never written
        const int ncol = min(p.dout - col0, NMAX);
        float w2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w2[j] = __ldg(p.dot_w + j);
        float* yd = p.Y + row * p.ldy + col0 / 16;
#pragma unroll
        for (int g = 0; g < NMAX / 16; ++g) {
          if (g * 16 < ncol) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = g * 16 + j;
              const float v = y[c] + ss->bias_s[c];
              a = fmaf(w2[j], 1.0f / (1.0f + expf(-v)), a);
            }
            yd[g] = a;
          }
        }
      } else if (!DOT && row < p.R) {
        float* yr = p.Y + (int64_t)blockIdx.z * p.split_stride + row * p.ldy + col0;
        const int ncol = min(p.dout - col0, NMAX);
        const bool vec = ((p.ldy & 3) == 0) && ((((uintptr_t)p.Y) & 15) == 0);
#pragma unroll
        for (int c4 = 0; c4 < NMAX / 4; ++c4) {
          if (c4 * 4 < ncol) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = c4 * 4 + i;
              float v = y[c] + ss->bias_s[c];
              o[i] = p.relu ? fmaxf(v, 0.f) : v;
            }
            if (vec && c4 * 4 + 3 < ncol) {
              *(float4*)(yr + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (c4 * 4 + i < ncol) yr[c4 * 4 + i] = o[i];
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 4 * NMAX);
  }
}


// split-K: Y[r, c] = act(sum_z part[z][r, c] + bias[c]) in a fixed order (deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t R, int dout,
                                     const float* __restrict__ bias, int relu, float* __restrict__ Y, int64_t ldy) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * dout) return;
  const int64_t r = i / dout;
  const int c = (int)(i % dout);
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += part[(int64_t)z * R * dout + i];
  if (bias) v += __ldg(bias + c);
  Y[r * ldy + c] = relu ? fmaxf(v, 0.f) : v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// fp32 [rows, cols] row-major with leading dimension ld; box = [KC, box_rows], SWIZZLE_128B;
// out-of-range elements (k >= cols, row >= rows) read as zero
static int make_tmap_f32(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld,
                         int box_rows) {
  EncodeTiledFn enc = encode_fn();
  B200_REQUIRE(enc, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace mlp
}  // namespace b200

extern "C" int64_t b200_linear_tf32x3_split_ld(int32_t din) { return ((int64_t)din + 3) / 4 * 4; }

extern "C" int b200_linear_tf32x3_split_weights(const float* Wt, int64_t ldw, int32_t din, int32_t dout,
                                                float* Wsplit, void* stream) {
  using namespace b200;
  using namespace b200::mlp;
  B200_REQUIRE(Wt && Wsplit && din > 0 && dout > 0 && ldw >= din, "bad arguments");
  const int64_t ld = b200_linear_tf32x3_split_ld(din);
  const int64_t n = (int64_t)dout * ld;
  split_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Wt, ldw, din, dout, ld, Wsplit);
  B200_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

static int launch_linear_tf32x3(const float* X, int64_t ldx, int64_t R, const float* Wt, int64_t ldw,
                                const float* Wsplit, const float* bias, int32_t din, int32_t dout,
                                int32_t relu, const float* dot_w, float* Y, int64_t ldy, void* stream,
                                int splits = 1, float* workspace = nullptr);

extern "C" int b200_linear_tf32x3(const float* X, int64_t ldx, int64_t R, const float* Wt, int64_t ldw,
                                  const float* Wsplit, const float* bias, int32_t din, int32_t dout,
                                  int32_t relu, float* Y, int64_t ldy, void* stream) {
  B200_REQUIRE(ldy >= dout, "leading dimension too small");
  return launch_linear_tf32x3(X, ldx, R, Wt, ldw, Wsplit, bias, din, dout, relu, nullptr, Y, ldy, stream);
}

extern "C" int b200_linear_tf32x3_sigmoid_dot(const float* X, int64_t ldx, int64_t R, const float* Wt, int64_t ldw,
                                              const float* Wsplit, const float* bias, int32_t din, int32_t dout,
                                              const float* dot_w16, float* A, int64_t lda, void* stream) {
  B200_REQUIRE(dot_w16 && A, "b200_linear_tf32x3_sigmoid_dot: null pointer");
  B200_REQUIRE(dout % 16 == 0 && lda >= dout / 16, "b200_linear_tf32x3_sigmoid_dot: dout must be a multiple of 16, lda >= dout / 16");
  return launch_linear_tf32x3(X, ldx, R, Wt, ldw, Wsplit, bias, din, dout, 0, dot_w16, A, lda, stream);
}

extern "C" int b200_linear_tf32x3_splitk(const float* X, int64_t ldx, int64_t R, const float* Wt, int64_t ldw,
                                         const float* bias, int32_t din, int32_t dout, int32_t relu, int32_t splits,
                                         float* workspace, size_t workspace_bytes, float* Y, int64_t ldy, void* stream) {
  B200_REQUIRE(splits >= 1 && splits <= 64, "b200_linear_tf32x3_splitk: splits outside [1, 64]");
  B200_REQUIRE(ldy >= dout, "leading dimension too small");
  B200_REQUIRE(splits == 1 || (workspace && workspace_bytes >= (size_t)splits * (size_t)R * (size_t)dout * 4),
               "b200_linear_tf32x3_splitk: workspace too small (splits * R * dout floats)");
  return launch_linear_tf32x3(X, ldx, R, Wt, ldw, nullptr, bias, din, dout, relu, nullptr, Y, ldy, stream, splits,
                              workspace);
}

static int launch_linear_tf32x3(const float* X, int64_t ldx, int64_t R, const float* Wt, int64_t ldw,
                                const float* Wsplit, const float* bias, int32_t din, int32_t dout,
                                int32_t relu, const float* dot_w, float* Y, int64_t ldy, void* stream,
                                int splits, float* workspace) {
  using namespace b200;
  using namespace b200::mlp;
  B200_REQUIRE(R >= 0 && din > 0 && dout > 0, "bad shape");
  B200_REQUIRE((ldx & 3) == 0 && ((uintptr_t)X & 15) == 0, "b200_linear_tf32x3 needs 16-byte aligned X rows (ldx % 4 == 0)");
  B200_REQUIRE(Wsplit ? (((uintptr_t)Wsplit & 15) == 0) : ((ldw & 3) == 0 && ((uintptr_t)Wt & 15) == 0),
               "b200_linear_tf32x3 needs 16-byte aligned weight rows (ldw % 4 == 0) or a split copy");
  B200_REQUIRE(ldx >= din && (Wsplit || ldw >= din), "leading dimension too small");
  if (R == 0) return 0;
  Params p;
  p.dot_w = dot_w;
  p.R = R;
  p.ldy = ldy;
  p.bias = bias;
  p.Y = Y;
  p.din = din;
  p.dout = dout;
  p.relu = relu;
  p.n_pad = dout >= NMAX ? NMAX : (dout + 31) / 32 * 32;
  p.n_tiles = (int)((R + TM - 1) / TM);
  p.n_chunks_total = (din + KC - 1) / KC;
  p.n_chunks = (p.n_chunks_total + splits - 1) / splits;
  splits = (p.n_chunks_total + p.n_chunks - 1) / p.n_chunks;        // no empty split
  p.split_stride = 0;
  if (splits > 1) {   // partial products [splits][R, dout] into the workspace, bias / ReLU in the reduction
    p.split_stride = R * (int64_t)dout;
    p.Y = workspace;
    p.ldy = dout;
    p.bias = nullptr;
    p.relu = 0;
  }
  const int stage_bytes = 2 * A_BYTES + 2 * p.n_pad * KC * 4;
  p.nstage = min(MAXSTAGE, (200 * 1024) / stage_bytes);
  const size_t smem = (size_t)p.nstage * stage_bytes + sizeof(Smem) + 1024;

  CUtensorMap tmX, tmW, tmWlo;
  if (make_tmap_f32(&tmX, X, R, din, ldx, TM)) return 1;
  if (Wsplit) {
    const int64_t ld = b200_linear_tf32x3_split_ld(din);
    if (make_tmap_f32(&tmW, Wsplit, dout, din, ld, p.n_pad)) return 1;
    if (make_tmap_f32(&tmWlo, Wsplit + (int64_t)dout * ld, dout, din, ld, p.n_pad)) return 1;
  } else {
    if (make_tmap_f32(&tmW, Wt, dout, din, ldw, p.n_pad)) return 1;
    tmWlo = tmW;
  }

  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(linear_tf32x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    B200_CUDA_OK(cudaFuncSetAttribute(linear_tf32x3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    B200_CUDA_OK(cudaFuncSetAttribute(linear_tf32x3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    B200_CUDA_OK(cudaFuncSetAttribute(linear_tf32x3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int gy = (dout + NMAX - 1) / NMAX;
  const int gx = max(1, min(p.n_tiles, sms / (gy * splits)));
  const dim3 grid(gx, gy, splits);
  cudaStream_t st = (cudaStream_t)stream;
  if (dot_w) {
    if (Wsplit) linear_tf32x3_kernel<true, true><<<grid, THREADS, smem, st>>>(tmX, tmW, tmWlo, p);
    else linear_tf32x3_kernel<false, true><<<grid, THREADS, smem, st>>>(tmX, tmW, tmWlo, p);
  } else {
    if (Wsplit) linear_tf32x3_kernel<true, false><<<grid, THREADS, smem, st>>>(tmX, tmW, tmWlo, p);
    else linear_tf32x3_kernel<false, false><<<grid, THREADS, smem, st>>>(tmX, tmW, tmWlo, p);
  }
  if (splits > 1) {
    const int64_t n = R * (int64_t)dout;
    splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, splits, R, dout, bias, relu, Y, ldy);
    count_launch();
  }
  B200_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
