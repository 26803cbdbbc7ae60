// ctc_beam.cu -- CTC prefix beam search on the device.
//
// Replaces torchaudio's `cuda_ctc_decoder(tokens, nbest=1, beam_size=10, blank_skip_threshold=0.95)` as the reference calls it at
// util/beam_infer.py:100-110 (once per exit, inference.py:66-79), for ALL exits and utterances of a forward in ONE launch.
// The library runs a handful of kernels per decoded frame per call (T' x 6 exits host-driven launches per batch); here one CTA owns
// one (exit, utterance) emission matrix and walks its T' frames on chip, so a whole (6, 64, 374, 256) forward decodes in a single
// launch with all 384 CTAs resident at once (256 threads, ~6 KB of shared memory each).
//
// Algorithm (prefix beam search without language model, merged prefixes, blank-frame skipping):
//   * frames t < enc_len with lp[t, blank] > log_blank_skip are not expanded: they count as a pure blank emission for every hypothesis
//     (p_b' = (p_b + p_nb) p(blank), p_nb' = 0) -- the behaviour of torchaudio's decoder that its outputs pin (oracle/ctc_beam_oracle.py);
//   * a hypothesis = (prefix, log p_blank, log p_non_blank); per frame every hypothesis i yields
//       stay      : p_b' = lse(p_b, p_nb) + lp[blank];  p_nb' = p_nb + lp[last_i]
//       extend c  : p_nb' = (c == last_i ? p_b : lse(p_b, p_nb)) + lp[c]          (a repeat extends only through a blank)
//     an extension that spells a prefix already in the beam is merged into that hypothesis' p_nb' (log-add) instead of competing;
//   * the `beam` best of the beam x V candidates by lse(p_b', p_nb') survive (ties: lower hypothesis index, then lower token id);
//   * prefixes are identified by (length, 64-bit rolling hash) and stored as a back-pointer trie in the caller's workspace.
// Thread c owns vocabulary entry c (and c + 256, ... for V > 256) for all hypotheses; the top-`beam` selection is `beam` rounds of a
// block-wide arg-max over per-thread register candidates.
#include "common.cuh"

namespace eec {
namespace {

constexpr int BS_THREADS = 256;
constexpr int BS_MAXBEAM = 16;
constexpr int BS_MAXVPT = 4;            // vocabulary entries per thread: V <= 1024
constexpr float BS_NEG = -INFINITY;

__device__ __forceinline__ float lse2(float a, float b) {
  if (a == BS_NEG) return b;
  if (b == BS_NEG) return a;
  const float m = fmaxf(a, b);
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ unsigned long long mix_hash(unsigned long long h, int c) {
  h = (h ^ ((unsigned long long)(c + 1) * 0x9E3779B97F4A7C15ull)) * 0xBF58476D1CE4E5B9ull;
  return h ^ (h >> 31);
}

struct BeamState {
  float pb[BS_MAXBEAM], pnb[BS_MAXBEAM];
  int last[BS_MAXBEAM], len[BS_MAXBEAM], node[BS_MAXBEAM];
  unsigned long long hash[BS_MAXBEAM];
};

template <int VPT>
__global__ void __launch_bounds__(BS_THREADS) ctc_beam_kernel(const float* __restrict__ lp, const int32_t* __restrict__ enc_len, int T, int V,
                                                              int beam, int nbest, int blank, float log_skip, int32_t* __restrict__ tokens,
                                                              int32_t* __restrict__ n_tokens, float* __restrict__ scores,
                                                              int32_t* __restrict__ ws) {
  pdl_trigger();
  pdl_wait();
  __shared__ BeamState st[2];
  __shared__ float tot[BS_MAXBEAM], stay_pb[BS_MAXBEAM], stay_pnb[BS_MAXBEAM];
  __shared__ unsigned killed[BS_MAXBEAM][BS_MAXVPT * BS_THREADS / 32];   // extension (i, c) merged into an existing hypothesis
  __shared__ float row[BS_MAXVPT * BS_THREADS];
  __shared__ float wbest_v[BS_THREADS / 32];
  __shared__ int wbest_i[BS_THREADS / 32];
  __shared__ float win_v[BS_MAXBEAM];
  __shared__ int win_i[BS_MAXBEAM];
  __shared__ int s_nb, s_nwin;

  const int utt = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ x = lp + (long)utt * T * V;
  int32_t* __restrict__ node_parent = ws + (long)utt * 2 * ((long)T * beam + 1);
  int32_t* __restrict__ node_tok = node_parent + ((long)T * beam + 1);
  const int n_frames = min(enc_len ? enc_len[utt] : T, T);

  if (tid == 0) {
    // the empty prefix: p_b = 1 (log 0), p_nb = 0
    st[0].pb[0] = 0.f; st[0].pnb[0] = BS_NEG; st[0].last[0] = -1; st[0].len[0] = 0; st[0].node[0] = -1;
    st[0].hash[0] = 0x243F6A8885A308D3ull;
    s_nb = 1;
  }
  int cur = 0, n_nodes = 0;
  float nxt[VPT];
#pragma unroll
  for (int k = 0; k < VPT; ++k) nxt[k] = (n_frames > 0 && tid + k * BS_THREADS < V) ? x[tid + k * BS_THREADS] : BS_NEG;
  __syncthreads();

  for (int t = 0; t < n_frames; ++t) {
    // this frame's emissions -> shared row; prefetch the next frame's into registers (the recursion is latency-bound)
#pragma unroll
    for (int k = 0; k < VPT; ++k) row[tid + k * BS_THREADS] = nxt[k];
    if (t + 1 < n_frames) {
#pragma unroll
      for (int k = 0; k < VPT; ++k)
        if (tid + k * BS_THREADS < V) nxt[k] = x[(long)(t + 1) * V + tid + k * BS_THREADS];
    }
    __syncthreads();
    const BeamState& S0 = st[cur];
    BeamState& S1 = st[cur ^ 1];
    const int nb = s_nb;
    if (row[blank] > log_skip) {
      // blank-dominated frame (uniform for the CTA): no expansion, no pruning -- but the frame is accounted for as a pure blank
      // emission, exactly like the library: p_b' = (p_b + p_nb) p(blank), p_nb' = 0 (a repeat after it starts a new token)
      if (tid < nb) {
        st[cur].pb[tid] = lse2(S0.pb[tid], S0.pnb[tid]) + row[blank];
        st[cur].pnb[tid] = BS_NEG;
      }
      __syncthreads();
      continue;
    }

    // A: stay probabilities of every hypothesis; clear the merge bitmaps
    if (tid < nb) {
      const float tt = lse2(S0.pb[tid], S0.pnb[tid]);
      tot[tid] = tt;
      stay_pb[tid] = tt + row[blank];
      stay_pnb[tid] = (S0.len[tid] > 0) ? S0.pnb[tid] + row[S0.last[tid]] : BS_NEG;
    }
    for (int k = tid; k < BS_MAXBEAM * (BS_MAXVPT * BS_THREADS / 32); k += BS_THREADS) (&killed[0][0])[k] = 0u;
    __syncthreads();
    // B: hypothesis j == hypothesis i extended by last_j  ->  that extension feeds j's p_nb' and leaves the candidate list
    if (tid < nb * nb) {
      const int j = tid / nb, i = tid - j * nb;
      if (i != j && S0.len[j] == S0.len[i] + 1 && S0.hash[j] == mix_hash(S0.hash[i], S0.last[j])) {
        const int c = S0.last[j];
        const float base = (S0.len[i] > 0 && c == S0.last[i]) ? S0.pb[i] : tot[i];
        stay_pnb[j] = lse2(stay_pnb[j], base + row[c]);       // (at most one i matches a given j: prefixes in the beam are distinct)
        atomicOr(&killed[i][c >> 5], 1u << (c & 31));
      }
    }
    __syncthreads();
    // C: this thread's candidates (hypothesis i, token c) in registers
    float cand[VPT][BS_MAXBEAM];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int c = tid + k * BS_THREADS;
#pragma unroll
      for (int i = 0; i < BS_MAXBEAM; ++i) {
        float v = BS_NEG;
        if (i < nb && c < V) {
          if (c == blank) v = lse2(stay_pb[i], stay_pnb[i]);
          else if (!((killed[i][c >> 5] >> (c & 31)) & 1u)) {
            const float base = (S0.len[i] > 0 && c == S0.last[i]) ? S0.pb[i] : tot[i];
            v = base + row[c];
          }
        }
        cand[k][i] = v;
      }
    }
    // D: `beam` rounds of block-wide arg-max (value desc, then flat index i * V + c asc)
    if (tid == 0) s_nwin = 0;
    for (int r = 0; r < beam; ++r) {
      float bv = BS_NEG;
      int bi = 0x7fffffff;
#pragma unroll
      for (int k = 0; k < VPT; ++k)
#pragma unroll
        for (int i = 0; i < BS_MAXBEAM; ++i) {
          const int idx = i * V + tid + k * BS_THREADS;
          if (cand[k][i] > bv || (cand[k][i] == bv && cand[k][i] != BS_NEG && idx < bi)) { bv = cand[k][i]; bi = idx; }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { wbest_v[warp] = bv; wbest_i[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        float v = lane < BS_THREADS / 32 ? wbest_v[lane] : BS_NEG;
        int ix = lane < BS_THREADS / 32 ? wbest_i[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
          if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
        }
        if (lane == 0) {
          win_v[r] = v; win_i[r] = ix;
          if (v != BS_NEG) s_nwin = r + 1;
        }
      }
      __syncthreads();
      const int wi = win_i[r];
      if (win_v[r] == BS_NEG) break;                               // fewer finite candidates than `beam` (uniform)
      const int wc = wi % V, wb = wi / V;
      if ((wc & (BS_THREADS - 1)) == tid) {                        // the owner retires the winner
#pragma unroll
        for (int k = 0; k < VPT; ++k)
#pragma unroll
          for (int i = 0; i < BS_MAXBEAM; ++i)
            if (k == wc / BS_THREADS && i == wb) cand[k][i] = BS_NEG;
      }
    }
    // E: the winners become the new beam (already in descending order of score)
    const int nwin = s_nwin;
    if (tid < nwin) {
      const int wi = win_i[tid], c = wi % V, i = wi / V;
      if (c == blank) {
        S1.pb[tid] = stay_pb[i]; S1.pnb[tid] = stay_pnb[i];
        S1.last[tid] = S0.last[i]; S1.len[tid] = S0.len[i]; S1.node[tid] = S0.node[i]; S1.hash[tid] = S0.hash[i];
      } else {
        const int nd = n_nodes + tid;
        node_parent[nd] = S0.node[i]; node_tok[nd] = c;
        S1.pb[tid] = BS_NEG; S1.pnb[tid] = win_v[tid];
        S1.last[tid] = c; S1.len[tid] = S0.len[i] + 1; S1.node[tid] = nd; S1.hash[tid] = mix_hash(S0.hash[i], c);
      }
    }
    if (tid == 0) s_nb = nwin;
    n_nodes += nwin;
    cur ^= 1;
    __syncthreads();
  }

  // results: hypotheses in beam order (= descending score), tokens by walking the back-pointer trie
  const BeamState& S = st[cur];
  const int nb = s_nb;
  if (tid < nbest) {
    int32_t* out = tokens + ((long)utt * nbest + tid) * T;
    if (tid < nb) {
      const int L = S.len[tid];
      int nd = S.node[tid];
      for (int k = L - 1; k >= 0; --k) { out[k] = node_tok[nd]; nd = node_parent[nd]; }
      for (int k = L; k < T; ++k) out[k] = -1;
      n_tokens[utt * nbest + tid] = L;
      scores[utt * nbest + tid] = lse2(S.pb[tid], S.pnb[tid]);
    } else {
      for (int k = 0; k < T; ++k) out[k] = -1;
      n_tokens[utt * nbest + tid] = 0;
      scores[utt * nbest + tid] = BS_NEG;
    }
  }
}

}  // namespace
}  // namespace eec

using namespace eec;

extern "C" int64_t eec_ctc_beam_workspace_bytes(int n_utt, int T, int beam) {
  return (int64_t)n_utt * 2 * ((int64_t)T * beam + 1) * (int64_t)sizeof(int32_t);
}

extern "C" int eec_ctc_beam_search(const float* lp, const int32_t* enc_len, int n_utt, int T, int V, int beam, int nbest, int blank,
                                   float log_blank_skip, int32_t* tokens, int32_t* n_tokens, float* scores, void* workspace,
                                   eec_stream_t stream) {
  EEC_CHECK_ARG(lp && tokens && n_tokens && scores && workspace, "ctc_beam_search: NULL argument");
  EEC_CHECK_ARG(V >= 2 && V <= BS_MAXVPT * BS_THREADS, "ctc_beam_search: vocabulary must be 2..%d (got %d)", BS_MAXVPT * BS_THREADS, V);
  EEC_CHECK_ARG(beam >= 1 && beam <= BS_MAXBEAM && beam <= V, "ctc_beam_search: beam must be 1..%d and <= V (got %d)", BS_MAXBEAM, beam);
  EEC_CHECK_ARG(nbest >= 1 && nbest <= beam, "ctc_beam_search: nbest must be 1..beam (got %d)", nbest);
  EEC_CHECK_ARG(blank >= 0 && blank < V, "ctc_beam_search: blank id out of range");
  if (n_utt == 0 || T == 0) return 0;
  int32_t* ws = reinterpret_cast<int32_t*>(workspace);
  const int vpt = cdiv(V, BS_THREADS);
#define EEC_BS_LAUNCH(VPT)                                                                                                   \
  launch_pdl(ctc_beam_kernel<VPT>, dim3(n_utt), dim3(BS_THREADS), 0, S(stream), lp, enc_len, T, V, beam, nbest, blank, log_blank_skip, \
             tokens, n_tokens, scores, ws)
  if (vpt == 1) EEC_BS_LAUNCH(1);
  else if (vpt == 2) EEC_BS_LAUNCH(2);
  else EEC_BS_LAUNCH(4);
#undef EEC_BS_LAUNCH
  EEC_LAUNCH_CHECK();
  return 0;
}
