// pf_attn.cu — masked joint text+video attention (placeholder kernel entry; host-side tile schedule is final).
#include <algorithm>
#include <vector>

#include "../../include/pf_b200.h"
#include "pf_common.cuh"

extern "C" int pf_attn_build_schedule(const int32_t* seg, const int32_t* time, int32_t batch, int32_t seq,
                                      int32_t* out, int64_t* allowed_pairs) {
  using namespace pf;
  PF_REQUIRE(batch > 0 && seq > 0, "pf_attn_build_schedule: bad shape");
  const int tiles = (seq + 127) / 128;
  const int stride = 1 + tiles;
  if (out == nullptr) return stride;
  PF_REQUIRE(seg && time, "pf_attn_build_schedule: null ids");
  for (int b = 0; b < batch; ++b) {
    const int32_t* sg = seg + static_cast<size_t>(b) * seq;
    const int32_t* tm = time + static_cast<size_t>(b) * seq;
    // per-tile summaries
    std::vector<int32_t> tmin(tiles), tmax(tiles), smin(tiles), smax(tiles);
    for (int t = 0; t < tiles; ++t) {
      const int lo = t * 128, hi = std::min(seq, lo + 128);
      int32_t a = tm[lo], bq = tm[lo], c = sg[lo], d = sg[lo];
      for (int i = lo; i < hi; ++i) {
        a = std::min(a, tm[i]);
        bq = std::max(bq, tm[i]);
        c = std::min(c, sg[i]);
        d = std::max(d, sg[i]);
      }
      tmin[t] = a; tmax[t] = bq; smin[t] = c; smax[t] = d;
    }
    int64_t pairs = 0;
    for (int qt = 0; qt < tiles; ++qt) {
      int32_t* row = out + (static_cast<size_t>(b) * tiles + qt) * stride;
      int cnt = 0;
      const int qlo = qt * 128, qhi = std::min(seq, qlo + 128);
      for (int kt = 0; kt < tiles; ++kt) {
        const int klo = kt * 128, khi = std::min(seq, klo + 128);
        // cheap rejections: no time-compatible pair, or disjoint segment ranges
        if (tmax[qt] < tmin[kt]) continue;
        if (smax[qt] < smin[kt] || smax[kt] < smin[qt]) continue;
        const bool uniform = (smin[qt] == smax[qt]) && (smin[kt] == smax[kt]) && (smin[qt] == smin[kt]);
        const bool full = uniform && (tmin[qt] >= tmax[kt]) && (khi - klo == 128);
        int64_t n_allowed = 0;
        if (full) {
          n_allowed = static_cast<int64_t>(qhi - qlo) * (khi - klo);
        } else {
          for (int i = qlo; i < qhi; ++i)
            for (int j = klo; j < khi; ++j) n_allowed += (sg[i] == sg[j] && tm[i] >= tm[j]) ? 1 : 0;
          if (n_allowed == 0) continue;
        }
        pairs += n_allowed;
        row[1 + cnt] = (kt << 1) | (full ? 0 : 1);
        ++cnt;
      }
      row[0] = cnt;
      for (int i = 1 + cnt; i < stride; ++i) row[i] = 0;
    }
    if (allowed_pairs) allowed_pairs[b] = pairs;
  }
  return stride;
}

extern "C" int pf_attn_fwd_masked(const pf_attn_desc* d, void* stream) {
  (void)d;
  (void)stream;
  pf::set_error("pf_attn_fwd_masked: kernel not built yet");
  return -1;
}
