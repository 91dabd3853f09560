// Pointwise kernels of the training backward (hold_b200/train_algo.py): everything between two products with
// a weight matrix (those run in k_mlp_tc<MLP_LINEAR>, hold_linear).  All operands are row-major fp32 [P, ld] matrices of <= 256
// columns; one thread per element, grid-stride.  Reference arithmetic: nn.Softplus(beta=100) with its threshold (shape_net.py:82),
// the Fourier embedding of engine/embedders.py:48-51 and its first / second derivative per component.
#pragma once
#include "common.cuh"
#include "mlp_simt.cuh"

namespace hold {

enum { EW_ACT = 0, EW_MUL = 1, EW_MULROW = 2, EW_U_DZ2 = 3, EW_DZ = 4, EW_EMBED = 5, EW_EMBED_VJP = 6, EW_EMBED_JVP = 7, EW_RELU = 8, EW_RELU_BWD = 9 };

struct EwArgs {
  const float* in0; const float* in1; const float* in2;
  float* out0; float* out1;
  int ld_in0, ld_in1, ld_in2, ld_out0, ld_out1;
  int ncols, aux;
};

// component e of the embedding of coordinate value t = x[e % 3]: value (order 0), d/dt (1), d2/dt2 (2)
__device__ __forceinline__ float embed_comp(int e, float t, int order) {
  if (e < 3) return order == 0 ? t : (order == 1 ? 1.f : 0.f);
  const int qq = (e - 3) / 3;
  const float f = (float)(1 << (qq >> 1));
  const float arg = t * f;
  const bool is_cos = (qq & 1) != 0;
  if (order == 0) return is_cos ? cosf(arg) : sinf(arg);
  if (order == 1) return is_cos ? -f * sinf(arg) : f * cosf(arg);
  return is_cos ? -f * f * cosf(arg) : -f * f * sinf(arg);
}

__global__ void __launch_bounds__(256) k_train_ew(int op, int P, EwArgs a) {
  // EW_ACT with a skip operand: `aux` appended columns (0 = the 39 embedding columns of the foreground SDF net)
  const size_t total = (size_t)P * (size_t)((op == EW_EMBED_VJP) ? 3 : ((op == EW_ACT && a.in1 != nullptr) ? a.ncols + (a.aux > 0 ? a.aux : kEmbed) : a.ncols));
  const int width = (int)(total / (size_t)P);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / width;
    const int c = (int)(i - p * width);
    switch (op) {
      case EW_ACT: {  // out0 = [softplus(z) | e], out1 = softplus'(z)
        if (c < a.ncols) {
          const float z = a.in0[p * a.ld_in0 + c];
          a.out0[p * a.ld_out0 + c] = softplus100(z);
          a.out1[p * a.ld_out1 + c] = softplus100_grad(z);
        } else {
          a.out0[p * a.ld_out0 + c] = a.in1[p * a.ld_in1 + c - a.ncols];
        }
      } break;
      case EW_MUL: a.out0[p * a.ld_out0 + c] = a.in0[p * a.ld_in0 + c] * a.in1[p * a.ld_in1 + c]; break;
      case EW_MULROW: a.out0[p * a.ld_out0 + c] = a.in0[c] * a.in1[p * a.ld_in1 + c]; break;
      case EW_U_DZ2: {  // u = h s; dz2 = h q softplus''(z), softplus'' = 100 s (1 - s); q may be one row (ld_in2 == 0)
        const float h = a.in0[p * a.ld_in0 + c], s = a.in1[p * a.ld_in1 + c], q = a.in2[p * a.ld_in2 + c];
        a.out0[p * a.ld_out0 + c] = h * s;
        a.out1[p * a.ld_out1 + c] = h * q * (100.0f * s * (1.0f - s));
      } break;
      case EW_DZ: {
        float v = a.in0[p * a.ld_in0 + c] * a.in1[p * a.ld_in1 + c];
        if (a.in2 != nullptr) v += a.in2[p * a.ld_in2 + c];
        a.out0[p * a.ld_out0 + c] = v;
      } break;
      case EW_EMBED: {  // in0 = x [P,3], in1 = embed weights [39] or NULL, aux = derivative order
        float v = (c < kEmbed) ? embed_comp(c, a.in0[p * a.ld_in0 + c % 3], a.aux) : 0.f;
        if (a.in1 != nullptr && c < kEmbed) v *= a.in1[c];
        a.out0[p * a.ld_out0 + c] = v;
      } break;
      case EW_EMBED_VJP: {  // out0[p, c] = sum over components e with e % 3 == c of d1[p, e] * ge[p, e]
        float acc = 0.f;
        for (int e = c; e < kEmbed; e += 3) acc += a.in0[p * a.ld_in0 + e] * a.in1[p * a.ld_in1 + e];
        a.out0[p * a.ld_out0 + c] = acc;
      } break;
      case EW_EMBED_JVP: a.out0[p * a.ld_out0 + c] = (c < kEmbed) ? a.in0[p * a.ld_in0 + c] * a.in1[p * a.ld_in1 + c % 3] : 0.f; break;
      case EW_RELU: a.out0[p * a.ld_out0 + c] = fmaxf(a.in0[p * a.ld_in0 + c], 0.f); break;
      case EW_RELU_BWD: a.out0[p * a.ld_out0 + c] = (a.in1[p * a.ld_in1 + c] > 0.f) ? a.in0[p * a.ld_in0 + c] : 0.f; break;
      default: break;
    }
  }
}

// Power-of-two scale of a matrix for the operand rescaling of hold_linear / hold_wgrad: 2^floor(log2(max |x|)) (1e-30 floor).
__global__ void __launch_bounds__(256) k_absmax_bits(int P, int ncols, const float* __restrict__ x, int ld, unsigned int* __restrict__ bits) {
  const size_t total = (size_t)P * ncols;
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / ncols;
    m = fmaxf(m, fabsf(x[p * ld + (i - p * ncols)]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}
__global__ void k_pow2_scale(const unsigned int* __restrict__ bits, float* __restrict__ out) {
  const float m = fmaxf(__uint_as_float(*bits), 1e-30f);
  int e;
  frexpf(m, &e);                 // m = f * 2^e, f in [0.5, 1)  ->  2^floor(log2 m) = 2^(e-1)
  *out = ldexpf(1.0f, e - 1);
}

}  // namespace hold
