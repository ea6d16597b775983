"""Result container of the decode API — field-for-field the reference's DecodeResult
(wenet/models/transformer/search.py:30-61)."""
from typing import List


class DecodeResult:

    def __init__(self,
                 tokens: List[int],
                 score: float = 0.0,
                 confidence: float = 0.0,
                 tokens_confidence: List[float] = None,
                 times: List[int] = None,
                 nbest: List[List[int]] = None,
                 nbest_scores: List[float] = None,
                 nbest_times: List[List[int]] = None,
                 text: str = ''):
        self.tokens = tokens
        self.score = score
        self.confidence = confidence
        self.tokens_confidence = tokens_confidence
        self.times = times
        self.nbest = nbest
        self.nbest_scores = nbest_scores
        self.nbest_times = nbest_times
        self.text = text


class LazyDecodeResult(DecodeResult):
    """DecodeResult whose token / time lists are materialised on first access.

    decode() returns one result per utterance with up to `beam` hypotheses of ~100 tokens each; building those as
    Python lists costs ~10 ms of interpreter time per 64-utterance batch (it dominated the host side and, with
    several batches in flight, the GIL).  The packed numpy rows are kept instead (`tok_rows`, `time_rows`
    [n_hyp, L] with `lens`), and `tokens / times / nbest / nbest_times / tokens_confidence / confidence` turn into the
    reference's plain Python objects (tuples / lists / floats) the first time they are read.
    """

    def __init__(self, score, nbest_scores, tok_rows, time_rows, lens, best=0, conf_fn=None, text=''):
        # deliberately no super().__init__: the lazy fields must stay absent from __dict__ until they are read
        self.score = score
        self.nbest_scores = nbest_scores
        self.text = text
        self._tok_rows, self._time_rows, self._lens, self._best, self._conf_fn = tok_rows, time_rows, lens, best, conf_fn

    def __getattr__(self, name):   # only reached for attributes that are not in __dict__ yet
        d = self.__dict__
        if name in ("confidence", "tokens_confidence"):
            fn = d.get("_conf_fn")
            conf, tc = fn() if fn is not None else (0.0, None)
            d["confidence"], d["tokens_confidence"] = conf, tc
            return d[name]
        if name.startswith("_"):
            raise AttributeError(name)
        lens = d["_lens"]
        if name == "nbest":
            v = [tuple(d["_tok_rows"][i, :int(lens[i])].tolist()) for i in range(len(lens))]
        elif name == "nbest_times":
            v = [d["_time_rows"][i, :int(lens[i])].tolist() for i in range(len(lens))]
        elif name == "tokens":
            b = d["_best"]
            v = tuple(d["_tok_rows"][b, :int(lens[b])].tolist())
        elif name == "times":
            b = d["_best"]
            v = d["_time_rows"][b, :int(lens[b])].tolist()
        else:
            raise AttributeError(name)
        d[name] = v
        return v

