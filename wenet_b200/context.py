"""Context-graph biasing for ctc_prefix_beam_search / attention_rescoring.

The reference builds an Aho-Corasick trie over the biasing phrases (`wenet/utils/context_graph.py:103-200`) and walks it
from Python inside the beam-search loop (`forward_one_step` :212-247, `finalize` :249-265, called at `search.py:171-173,
200-203, 229-234`).  Here the graph is flattened once into arrays (children in CSR form sorted by token, fail arcs, the
three per-node scores as IEEE doubles - the reference's Python floats) and the walk runs inside the CUDA beam-search
kernel (csrc/search.cu) with the same arithmetic in the same order.

`flatten(graph)` accepts the reference's own `ContextGraph` object (anything with a `.root` whose nodes carry
`.next / .fail / .token / .token_score / .node_score / .output_score`) - `recognize.py --context_bias_mode` keeps
building it with the reference's code - or a list of token-id lists via `build(phrases, context_score)`.
"""
from collections import deque
from typing import Dict, List, Sequence

import numpy as np


class ContextArrays:
    """Flattened context graph; node 0 is the root (token -1)."""
    __slots__ = ("child_off", "child_tok", "child_node", "fail", "token", "node_score", "token_score", "output_score",
                 "_dev")

    def __init__(self, child_off, child_tok, child_node, fail, token, node_score, token_score, output_score):
        self.child_off = np.ascontiguousarray(child_off, dtype=np.int32)
        self.child_tok = np.ascontiguousarray(child_tok, dtype=np.int32)
        self.child_node = np.ascontiguousarray(child_node, dtype=np.int32)
        self.fail = np.ascontiguousarray(fail, dtype=np.int32)
        self.token = np.ascontiguousarray(token, dtype=np.int32)
        self.node_score = np.ascontiguousarray(node_score, dtype=np.float64)
        self.token_score = np.ascontiguousarray(token_score, dtype=np.float64)
        self.output_score = np.ascontiguousarray(output_score, dtype=np.float64)
        self._dev = {}

    @property
    def num_nodes(self) -> int:
        return int(self.fail.shape[0])

    def child(self, node: int, token: int) -> int:
        lo, hi = int(self.child_off[node]), int(self.child_off[node + 1])
        i = lo + int(np.searchsorted(self.child_tok[lo:hi], token))
        return int(self.child_node[i]) if i < hi and int(self.child_tok[i]) == token else -1


def flatten(graph) -> ContextArrays:
    """Breadth-first numbering of a reference-style graph (root first)."""
    if isinstance(graph, ContextArrays):
        return graph
    root = graph.root
    order, index = [root], {id(root): 0}
    q = deque([root])
    while q:
        n = q.popleft()
        for tok in sorted(n.next):
            c = n.next[tok]
            if id(c) not in index:
                index[id(c)] = len(order)
                order.append(c)
                q.append(c)
    off, ctok, cnode = [0], [], []
    for n in order:
        for tok in sorted(n.next):
            ctok.append(int(tok))
            cnode.append(index[id(n.next[tok])])
        off.append(len(ctok))
    fail = [index[id(n.fail)] for n in order]
    return ContextArrays(off, ctok, cnode, fail, [int(n.token) for n in order], [float(n.node_score) for n in order],
                         [float(n.token_score) for n in order], [float(n.output_score) for n in order])


class _Node:
    def __init__(self, token, token_score, node_score, output_score, is_end):
        self.token, self.token_score, self.node_score = token, token_score, node_score
        self.output_score, self.is_end = output_score, is_end
        self.next: Dict[int, "_Node"] = {}
        self.fail = None
        self.output = None


class _Graph:
    pass


def build(phrases: Sequence[Sequence[int]], context_score: float = 6.0) -> ContextArrays:
    """The reference's construction (context_graph.py:126-200: trie, then fail / output arcs by BFS) from token-id
    lists, for callers that have no `wenet` import at hand; `flatten()` of the reference's own object is equivalent."""
    root = _Node(-1, 0, 0, 0, False)
    root.fail = root
    for tokens in phrases:
        node = root
        for i, tok in enumerate(tokens):
            if tok not in node.next:
                is_end = i == len(tokens) - 1
                ns = node.node_score + context_score
                node.next[tok] = _Node(tok, context_score, ns, ns if is_end else 0, is_end)
            node = node.next[tok]
    q = deque()
    for tok, node in root.next.items():
        node.fail = root
        q.append(node)
    while q:
        cur = q.popleft()
        for tok, node in cur.next.items():
            fail = cur.fail
            if tok in fail.next:
                fail = fail.next[tok]
            else:
                fail = fail.fail
                while tok not in fail.next:
                    fail = fail.fail
                    if fail.token == -1:
                        break
                if tok in fail.next:
                    fail = fail.next[tok]
            node.fail = fail
            output = node.fail
            while not output.is_end:
                output = output.fail
                if output.token == -1:
                    output = None
                    break
            node.output = output
            node.output_score += 0 if output is None else output.output_score
            q.append(node)
    g = _Graph()
    g.root = root
    return flatten(g)


def to_device(arr: ContextArrays, device):
    """torch tensors of the arrays on `device` (cached per device) + the ctypes struct the C ABI takes."""
    import ctypes as C

    import torch

    from ._lib import WbContextGraph
    key = str(device)
    if key not in arr._dev:
        t = {n: torch.from_numpy(getattr(arr, n)).to(device) for n in
             ("child_off", "child_tok", "child_node", "fail", "token", "node_score", "token_score", "output_score")}
        s = WbContextGraph(num_nodes=arr.num_nodes,
                           child_off=C.c_void_p(t["child_off"].data_ptr()), child_tok=C.c_void_p(t["child_tok"].data_ptr()),
                           child_node=C.c_void_p(t["child_node"].data_ptr()), fail=C.c_void_p(t["fail"].data_ptr()),
                           token=C.c_void_p(t["token"].data_ptr()), node_score=C.c_void_p(t["node_score"].data_ptr()),
                           token_score=C.c_void_p(t["token_score"].data_ptr()),
                           output_score=C.c_void_p(t["output_score"].data_ptr()))
        arr._dev[key] = (t, s)
    return arr._dev[key][1]
