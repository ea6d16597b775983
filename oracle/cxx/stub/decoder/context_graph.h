// TEST INFRASTRUCTURE ONLY.  Stand-in for runtime/core/decoder/context_graph.h, whose real version needs openfst
// (fst/compose.h, fst/vector-fst.h; not installed).  The reference's CtcPrefixBeamSearch only ever calls
// ContextGraph::GetNextState, and only when a graph was passed to its constructor; oracle/_ref/ctc_search_ref never passes
// one, so this class is never instantiated - it exists so that decoder/ctc_prefix_beam_search.{h,cc} compile unmodified.
#ifndef ORACLE_STUB_DECODER_CONTEXT_GRAPH_H_
#define ORACLE_STUB_DECODER_CONTEXT_GRAPH_H_
namespace wenet {
class ContextGraph {
 public:
  int GetNextState(int /*cur_state*/, int /*word_id*/, float* score) {
    *score = 0.0f;
    return 0;
  }
};
}  // namespace wenet
#endif  // ORACLE_STUB_DECODER_CONTEXT_GRAPH_H_
