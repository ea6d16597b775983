"""Write a `.wbm` model file for the C++ runtime back-end (runtime/b200_asr_model.h): the wb_model_config struct, sos /
eos / bidirectional flag and the tensors `weights.pack_state_dict` produces for `wb_model_set_tensor`, in one flat file.

    python -m wenet_b200.export train.yaml final.pt model.wbm [--precise]
"""
import struct
import sys
from typing import Dict

import numpy as np
import torch

from ._lib import WbModelConfig
from .weights import WB_BF16, WB_F32, ModelSpec, pack_state_dict

MAGIC = b"WBM0001\0"


def export_model(configs: dict, state_dict: Dict[str, torch.Tensor], path: str, precise: bool = False,
                 with_decoder: bool = True) -> int:
    spec = ModelSpec(configs)
    has_dec = with_decoder and any(k.startswith("decoder.") for k in state_dict)
    cfg = WbModelConfig(
        input_dim=spec.input_dim, d_model=spec.d_model, heads=spec.heads, ffn_dim=spec.ffn_dim,
        enc_layers=spec.enc_layers, cnn_kernel=spec.cnn_kernel, cnn_causal=int(spec.cnn_causal),
        cnn_norm=0 if spec.cnn_norm == "layer_norm" else 1, vocab=spec.vocab,
        dec_layers=spec.dec_layers if has_dec else 0, rdec_layers=spec.rdec_layers if has_dec else 0,
        dec_heads=spec.dec_heads, dec_ffn_dim=spec.dec_ffn_dim, max_pos=spec.max_pos, has_cmvn=int(spec.has_cmvn),
        precise=int(precise), ln_eps=spec.ln_eps, dec_ln_eps=spec.dec_ln_eps,
        arch=int(spec.arch), dec_flavor=int(spec.dec_flavor), dec_max_len=int(spec.dec_max_len))
    packed = {k: v for k, v in pack_state_dict(spec, state_dict, precise=precise).items()
              if has_dec or not k.startswith("dec.")}
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(bytes(cfg))
        f.write(struct.pack("<4i", spec.sos, spec.eos, int(spec.bidirectional and has_dec), len(packed)))
        for name, t in packed.items():
            if t.dtype == torch.bfloat16:
                arr, dt = t.view(torch.int16).numpy(), WB_BF16
            else:
                arr, dt = t.numpy(), WB_F32
            arr = np.ascontiguousarray(arr)
            nb = name.encode()
            f.write(struct.pack("<i", len(nb)))
            f.write(nb)
            f.write(struct.pack("<iq", dt, arr.size))
            f.write(arr.tobytes())
    return len(packed)


if __name__ == "__main__":
    import yaml
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if len(args) != 3:
        raise SystemExit(__doc__)
    with open(args[0]) as fin:
        cfg = yaml.safe_load(fin)
    sd = torch.load(args[1], map_location="cpu")
    n = export_model(cfg, sd, args[2], precise="--precise" in sys.argv)
    print("wrote %d tensors to %s" % (n, args[2]))
