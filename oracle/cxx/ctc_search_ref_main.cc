// TEST INFRASTRUCTURE ONLY (oracle/_ref/ctc_search_ref): the REFERENCE's C++ CTC prefix beam search
// (/root/reference/runtime/core/decoder/ctc_prefix_beam_search.{h,cc} + utils/utils.{h,cc}, compiled from where they lie)
// behind a stdin/stdout driver.  Input (text): "T V beam" followed by T * V log-probabilities, any number of such blocks.
// Output per block: "N" then N lines "score viterbi_score | tok tok ... | time time ...", best first - what
// CtcPrefixBeamSearch::Inputs() / Likelihood() / Times() hold after Search() + FinalizeSearch()
// (the same sequence as runtime/core/test/ctc_prefix_beam_search_test.cc:29-72).
#include <cstdio>
#include <vector>

#include "decoder/ctc_prefix_beam_search.h"

int main() {
  int T, V, beam;
  while (std::scanf("%d %d %d", &T, &V, &beam) == 3) {
    std::vector<std::vector<float>> logp(T, std::vector<float>(V));
    for (int t = 0; t < T; ++t)
      for (int v = 0; v < V; ++v)
        if (std::scanf("%f", &logp[t][v]) != 1) return 1;
    wenet::CtcPrefixBeamSearchOptions opts;
    opts.blank = 0;
    opts.first_beam_size = beam;
    opts.second_beam_size = beam;
    wenet::CtcPrefixBeamSearch search(opts);
    search.Search(logp);
    search.FinalizeSearch();
    const auto& hyps = search.Inputs();
    const auto& like = search.Likelihood();
    const auto& vit = search.viterbi_likelihood();
    const auto& times = search.Times();
    std::printf("%zu\n", hyps.size());
    for (size_t i = 0; i < hyps.size(); ++i) {
      std::printf("%.9g %.9g |", like[i], i < vit.size() ? vit[i] : 0.0f);
      for (int tok : hyps[i]) std::printf(" %d", tok);
      std::printf(" |");
      for (int tm : times[i]) std::printf(" %d", tm);
      std::printf("\n");
    }
  }
  return 0;
}
