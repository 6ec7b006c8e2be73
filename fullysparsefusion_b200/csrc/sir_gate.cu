// sir_gate.cu — SIRLayer's relative-position gate, fully fused (a12).
//
// Reference: SIRLayer.forward of the un-vendored mmdet3d fork, built by SIR at
// projects/mmdet3d_plugin/models/backbones/sir.py:41-62 (rel_mlp = build_mlp(3, [16, 32, Cin], LN, act),
// rel_dist_scaler = 10, xyz_normalizer = [20, 20, 4]):
//     x    = cat(features[:, :3] / xyz_normalizer, features[:, 3:])
//     gate = rel_mlp(f_cluster / rel_dist_scaler)           # 3 → h1 → h2 → Cin, Linear(bias=False) → LN → act each
//     out  = x * gate
// As three GEMM launches the tiny K (3, 16, 32) leaves the tensor cores idle and the [n, Cin] gate makes two
// extra trips through HBM; here one warp walks points, keeps the three weight matrices in shared memory
// (~25 KB), does the 3 → 16 → 32 → Cin chain with fp32 FMAs + warp-shuffle LayerNorms, and writes x * gate
// directly.  HBM-bound: algorithmic bytes = 4·n·(2·Cin + 3).
#include "gemm_common.cuh"

namespace fsfb {

constexpr int kGateMaxH = 32;    // hidden widths h1, h2 <= 32 (the stock [16, 32])
constexpr int kGateMaxC = 256;   // Cin <= 256 (180 / 136 / 133 in the stock configs)
constexpr int kGateWarps = 8;

struct GateParams {
  const float* feats;
  int64_t n;
  int c;
  int64_t feat_stride;
  const float* feats_b;   // optional second source: channels [c_a, c) come from feats_b[:, 0:c-c_a]
  int64_t feats_b_stride;
  int c_a;
  const float* f_cluster;
  int64_t fc_stride;
  float inv_dummy;
  float scaler;       // f_cluster / scaler
  float nrm[3];       // xyz_normalizer
  int h1, h2;
  const float *w1, *g1, *b1;  // [h1,3], LN(h1)
  const float *w2, *g2, *b2;  // [h2,h1], LN(h2)
  const float *w3, *g3, *b3;  // [c,h2], LN(c)
  float eps;
  int act;
  float* out;
  int64_t out_stride;
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NT>   // 32-channel groups per lane: ceil(c / 32), so that no unrolled slot is a predicated-off instruction
__global__ void __launch_bounds__(kGateWarps * 32) k_sir_gate(const GateParams P) {
  extern __shared__ float s_w[];
  // layout: w1t [3][32] | w2t [32][32] | w3t [32][cpad] | g1,b1 [32] | g2,b2 [32] | g3,b3 [cpad]
  const int cpad = (P.c + 31) & ~31;
  float* w1t = s_w;
  float* w2t = w1t + 3 * 32;
  float* w3t = w2t + 32 * 32;
  float* g1 = w3t + 32 * cpad;
  float* b1 = g1 + 32;
  float* g2 = b1 + 32;
  float* b2 = g2 + 32;
  float* g3 = b2 + 32;
  float* b3 = g3 + cpad;
  float4* hx = reinterpret_cast<float4*>(b3 + cpad) + (threadIdx.x >> 5) * 32;   // this warp's [32 hidden units] x [4 points]
  for (int t = threadIdx.x; t < 3 * 32; t += blockDim.x) {
    const int i = t / 32, j = t % 32;
    w1t[t] = j < P.h1 ? __ldg(P.w1 + j * 3 + i) : 0.f;
  }
  for (int t = threadIdx.x; t < 32 * 32; t += blockDim.x) {
    const int i = t / 32, j = t % 32;
    w2t[t] = (j < P.h2 && i < P.h1) ? __ldg(P.w2 + j * P.h1 + i) : 0.f;
  }
  for (int t = threadIdx.x; t < 32 * cpad; t += blockDim.x) {
    const int i = t / cpad, j = t % cpad;
    w3t[t] = (j < P.c && i < P.h2) ? __ldg(P.w3 + (int64_t)j * P.h2 + i) : 0.f;
  }
  for (int t = threadIdx.x; t < 32; t += blockDim.x) {
    g1[t] = t < P.h1 ? __ldg(P.g1 + t) : 0.f;
    b1[t] = t < P.h1 ? __ldg(P.b1 + t) : 0.f;
    g2[t] = t < P.h2 ? __ldg(P.g2 + t) : 0.f;
    b2[t] = t < P.h2 ? __ldg(P.b2 + t) : 0.f;
  }
  for (int t = threadIdx.x; t < cpad; t += blockDim.x) {
    g3[t] = t < P.c ? __ldg(P.g3 + t) : 0.f;
    b3[t] = t < P.c ? __ldg(P.b3 + t) : 0.f;
  }
  __syncthreads();

  const int lane = lane_id();
  constexpr int PT = 4;      // points in flight per warp: independent dependency chains for ILP (hx holds float4 = PT values)
  const int64_t warps = (int64_t)gridDim.x * kGateWarps;
  for (int64_t i0 = ((int64_t)blockIdx.x * kGateWarps + (threadIdx.x >> 5)) * PT; i0 < P.n; i0 += warps * PT) {
    // the gated rows themselves: requested first, consumed last (their latency hides behind the three layers)
    float xin[PT][NT];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      const int64_t i = min(i0 + p, P.n - 1);
      const float* fr = P.feats + i * P.feat_stride;
      const float* fb = P.feats_b ? P.feats_b + i * P.feats_b_stride - P.c_a : fr;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int c = lane + 32 * t;
        xin[p][t] = c < P.c ? __ldg((c < P.c_a ? fr : fb) + c) : 0.f;
      }
    }
    // ---- layer 1: 3 → h1, LN, act (lane j owns hidden unit j) ----
    float h[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      const int64_t i = min(i0 + p, P.n - 1);
      const float* fc = P.f_cluster + i * P.fc_stride;
      const float fx = __fdiv_rn(__ldg(fc), P.scaler), fy = __fdiv_rn(__ldg(fc + 1), P.scaler), fz = __fdiv_rn(__ldg(fc + 2), P.scaler);
      h[p] = fmaf(fz, w1t[2 * 32 + lane], fmaf(fy, w1t[32 + lane], fx * w1t[lane]));
    }
    {
      const bool on = lane < P.h1;
      float mu[PT], d[PT], var[PT];
#pragma unroll
      for (int p = 0; p < PT; ++p) mu[p] = on ? h[p] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int p = 0; p < PT; ++p) mu[p] += __shfl_xor_sync(0xffffffffu, mu[p], o);
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        d[p] = on ? h[p] - mu[p] / (float)P.h1 : 0.f;
        var[p] = d[p] * d[p];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int p = 0; p < PT; ++p) var[p] += __shfl_xor_sync(0xffffffffu, var[p], o);
#pragma unroll
      for (int p = 0; p < PT; ++p)
        h[p] = on ? apply_act(d[p] * (1.f / sqrtf(var[p] / (float)P.h1 + P.eps)) * g1[lane] + b1[lane], P.act) : 0.f;
    }
    // ---- layer 2: h1 → h2 ----
    float h2[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) h2[p] = 0.f;
    __syncwarp();
    hx[lane] = make_float4(h[0], h[1], h[2], h[3]);   // lanes >= h1 hold zeros
    __syncwarp();
#pragma unroll 8
    for (int k = 0; k < kGateMaxH; ++k) {
      if (k < P.h1) {
        const float w = w2t[k * 32 + lane];
        const float4 hv = hx[k];   // broadcast: one shared-memory read instead of four shuffles
        h2[0] = fmaf(hv.x, w, h2[0]); h2[1] = fmaf(hv.y, w, h2[1]); h2[2] = fmaf(hv.z, w, h2[2]); h2[3] = fmaf(hv.w, w, h2[3]);
      }
    }
    {
      const bool on = lane < P.h2;
      float mu[PT], d[PT], var[PT];
#pragma unroll
      for (int p = 0; p < PT; ++p) mu[p] = on ? h2[p] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int p = 0; p < PT; ++p) mu[p] += __shfl_xor_sync(0xffffffffu, mu[p], o);
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        d[p] = on ? h2[p] - mu[p] / (float)P.h2 : 0.f;
        var[p] = d[p] * d[p];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int p = 0; p < PT; ++p) var[p] += __shfl_xor_sync(0xffffffffu, var[p], o);
#pragma unroll
      for (int p = 0; p < PT; ++p)
        h2[p] = on ? apply_act(d[p] * (1.f / sqrtf(var[p] / (float)P.h2 + P.eps)) * g2[lane] + b2[lane], P.act) : 0.f;
    }
    // ---- layer 3: h2 → c (lane owns channels lane + 32 t) ----
    float g[PT][NT];
#pragma unroll
    for (int p = 0; p < PT; ++p)
#pragma unroll
      for (int t = 0; t < NT; ++t) g[p][t] = 0.f;
    __syncwarp();
    hx[lane] = make_float4(h2[0], h2[1], h2[2], h2[3]);
    __syncwarp();
#pragma unroll 2
    for (int k = 0; k < kGateMaxH; ++k) {
      if (k < P.h2) {
        const float4 hq = hx[k];
        const float hv[PT] = {hq.x, hq.y, hq.z, hq.w};
        const float* wr = w3t + k * cpad + lane;
#pragma unroll
        for (int t = 0; t < NT; ++t)
          {
            const float w = wr[32 * t];
#pragma unroll
            for (int p = 0; p < PT; ++p) g[p][t] = fmaf(hv[p], w, g[p][t]);
          }
      }
    }
    float mu[PT], q[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      mu[p] = 0.f;
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (lane + 32 * t < P.c) mu[p] += g[p][t];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int p = 0; p < PT; ++p) mu[p] += __shfl_xor_sync(0xffffffffu, mu[p], o);
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      mu[p] /= (float)P.c;
      q[p] = 0.f;
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (lane + 32 * t < P.c) {
          const float d = g[p][t] - mu[p];
          q[p] += d * d;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int p = 0; p < PT; ++p) q[p] += __shfl_xor_sync(0xffffffffu, q[p], o);
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      const int64_t i = i0 + p;
      if (i >= P.n) break;
      const float rstd = 1.f / sqrtf(q[p] / (float)P.c + P.eps);
      float* o = P.out + i * P.out_stride;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int c = lane + 32 * t;
        if (c < P.c) {
          const float gate = apply_act((g[p][t] - mu[p]) * rstd * g3[c] + b3[c], P.act);
          float x = xin[p][t];
          if (c < 3) x = __fdiv_rn(x, P.nrm[c]);
          o[c] = __fmul_rn(x, gate);
        }
      }
    }
  }
}

}  // namespace fsfb

extern "C" int fsfb_sir_gate_input(const float* feats, int64_t n, int c, int64_t feat_stride, const float* feats_b,
                                   int64_t feats_b_stride, int c_a, const float* f_cluster,
                                   int64_t fc_stride, float rel_dist_scaler, const float* xyz_normalizer, int h1, int h2,
                                   const float* w1, const float* ln1_w, const float* ln1_b, const float* w2,
                                   const float* ln2_w, const float* ln2_b, const float* w3, const float* ln3_w,
                                   const float* ln3_b, float eps, int act, float* out, int64_t out_stride, void* stream) {
  using namespace fsfb;
  if (!feats_b) c_a = c;
  FSFB_CHECK_ARG(n >= 0 && c >= 3 && c <= kGateMaxC && c_a >= 3 && c_a <= c && feat_stride >= c_a && out_stride >= c &&
                     fc_stride >= 3 && (!feats_b || feats_b_stride >= c - c_a),
                 "sir_gate_input: bad shape (c must be 3..%d)", kGateMaxC);
  FSFB_CHECK_ARG(h1 >= 1 && h1 <= kGateMaxH && h2 >= 1 && h2 <= kGateMaxH, "sir_gate_input: hidden widths must be 1..%d", kGateMaxH);
  FSFB_CHECK_ARG(act == FSFB_ACT_NONE || act == FSFB_ACT_RELU || act == FSFB_ACT_GELU, "sir_gate_input: bad act");
  FSFB_CHECK_ARG(xyz_normalizer && rel_dist_scaler != 0.f, "sir_gate_input: bad normalizers");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(feats && f_cluster && out && w1 && ln1_w && ln1_b && w2 && ln2_w && ln2_b && w3 && ln3_w && ln3_b,
                 "sir_gate_input: null pointer");
  GateParams P;
  P.feats = feats; P.n = n; P.c = c; P.feat_stride = feat_stride; P.feats_b = feats_b; P.feats_b_stride = feats_b_stride; P.c_a = c_a; P.f_cluster = f_cluster; P.fc_stride = fc_stride;
  P.inv_dummy = 0.f; P.scaler = rel_dist_scaler;
  P.nrm[0] = xyz_normalizer[0]; P.nrm[1] = xyz_normalizer[1]; P.nrm[2] = xyz_normalizer[2];
  P.h1 = h1; P.h2 = h2;
  P.w1 = w1; P.g1 = ln1_w; P.b1 = ln1_b; P.w2 = w2; P.g2 = ln2_w; P.b2 = ln2_b; P.w3 = w3; P.g3 = ln3_w; P.b3 = ln3_b;
  P.eps = eps; P.act = act; P.out = out; P.out_stride = out_stride;
  const int cpad = (c + 31) & ~31;
  const size_t smem = (size_t)(3 * 32 + 32 * 32 + 32 * cpad + 4 * 32 + 2 * cpad + kGateWarps * 32 * 4) * 4;
  const int grid = (int)std::min<int64_t>(ceil_div(n, kGateWarps * 4), (int64_t)kNumSMs * 3);  // 3 CTAs of 256 threads fit an SM
  auto go = [&](auto kern) -> int {
    static bool attr = false;   // one per instantiation of this generic lambda
    if (!attr) {
      FSFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr = true;
    }
    FSFB_LAUNCH(kern, grid, kGateWarps * 32, smem, (cudaStream_t)stream, P);
    return FSFB_OK;
  };
  switch (cpad / 32) {
    case 1: return go(k_sir_gate<1>);
    case 2: return go(k_sir_gate<2>);
    case 3: return go(k_sir_gate<3>);
    case 4: return go(k_sir_gate<4>);
    case 5: return go(k_sir_gate<5>);
    case 6: return go(k_sir_gate<6>);
    case 7: return go(k_sir_gate<7>);
    default: return go(k_sir_gate<8>);
  }
}
