"""Utterance-level data parallelism: every stage of the hot path is
per-utterance, so a batch shards across GPUs with NO data-path collective — each rank decodes its
own utterances with a full weight replica; only token lists (kilobytes) are gathered on the host.

Partitioning: longest-processing-time-first greedy on an attention-aware cost T'^2 * c1 + T' * c2
(the reference decodes padded batches in list order; recognize.py sorts nothing across processes)."""
from typing import List, Sequence


def utterance_cost(num_frames: int, d_model: int = 256) -> float:
    tp = max(((num_frames - 1) // 2 - 1) // 2, 0)
    # per-layer FLOPs: linear part ~ 2*T'*d*(2*8d + 4d + 3d) ; attention part ~ 4*T'^2*d
    return tp * d_model * 46.0 * d_model + 4.0 * tp * tp * d_model


def shard_utterances(num_frames: Sequence[int], world_size: int, rank: int, d_model: int = 256) -> List[int]:
    """Indices (into the global list) of the utterances rank `rank` processes.  Deterministic and
    identical on every rank; the union over ranks is a partition of range(len(num_frames))."""
    assert 0 <= rank < world_size
    order = sorted(range(len(num_frames)), key=lambda i: (-utterance_cost(num_frames[i], d_model), i))
    loads = [0.0] * world_size
    owner = [0] * len(num_frames)
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        owner[i] = r
        loads[r] += utterance_cost(num_frames[i], d_model)
    return [i for i in range(len(num_frames)) if owner[i] == rank]
