// Host interface of the fused transformer stack (token_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace ofb {

// one linear layer of a Transformer_Block in the split-half weight format of the tcgen05 engine:
// ws = [hi plane (cout, cin) ; lo plane (cout, cin)] fp16 of weight / unscale, bias (cout) float32
struct TokLinear { const void* ws; const float* bias; float unscale; };
struct TokBlockDesc {
  TokLinear lin[4];                       // fused q|kv projection (1536 x 512), attn.proj (512 x 512), mlp.fc1 (2048 x 512), mlp.fc2 (512 x 2048)
  const float *n1g, *n1b, *n2g, *n2b;     // norm1 / norm2 affine parameters (512)
};

constexpr int kTokMaxBlocks = 6;
constexpr int kTokMaxPanos = 4096;      // panoramas per launch
constexpr int kTokLanes = 2;            // concurrent launches of one handle (engine lanes)

// kernel argument (one __grid_constant__ block): per transformer block four weight tensor maps + epilogue constants
struct TokBlock {
  CUtensorMap w[4];
  const float* bias[4];
  float unscale[4];
  const float *n1g, *n1b, *n2g, *n2b;
};
struct TokStack {
  TokBlock blk[kTokMaxBlocks];
  const float* enc_g;
  const float* enc_b;
  unsigned int* counters;                // device: [lane][panorama][2] arrival counters, zero between launches
  int nblk;
};

bool token_stack_supported(int N);                       // tokens per panorama the kernel handles
size_t token_stack_scratch_floats(int B, int N);          // exchange buffers (scores, attention output, fc2 partial sums)
int token_stack_prepare(const TokBlockDesc* blocks, int nblk, const float* enc_g, const float* enc_b, TokStack* out);
// x: (B, N, 512) float32 residual stream, updated in place; enc_out: (B, N, 512) float32 = encoder_norm(x) (written
// when the whole stack runs: stop_phase = 0); stop_phase > 0 runs only the first stop_phase GEMM phases (4 per block;
// tests look at the exchange buffers)
int token_stack_launch(const TokStack& st, float* x, float* scratch, size_t scratch_floats, float* enc_out, int B, int N,
                       int stop_phase, int lane, bool pdl, cudaStream_t s);
void token_stack_release(TokStack* st);                  // frees the counters
int token_stack_resident_groups(int N);                  // panoramas (groups of 16 CTAs) the device runs at once

}  // namespace ofb
