"""Generates tests/golden/*.pt by running the UNMODIFIED reference modules from /root/reference
(build container only - the GPU box has no /root/reference, it only reads the fixtures).

    python oracle/make_golden.py

The reference's TextEncoder downloads the BioBERT checkpoint in __init__ (src/text_encoder.py:46-47);
offline that raises OSError, so AutoConfig/AutoModel.from_pretrained are replaced by a random-init
BertModel(BertConfig(BIOBERT_BASE)) for the construction only, and cnn_encoder.pretrained=False stops
the ImageNet download (src/cnn_encoder.py:76-77).  Everything executed afterwards is reference code.
The weights come from the drop-in module built under a fixed seed and are loaded into the reference
model with load_state_dict(strict=True) - which also proves the state_dict contract.
"""

from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("MRD_REFERENCE", "/root/reference")

import synth  # noqa: E402


def build_reference():
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        from transformers import BertConfig, BertModel

        import src.text_encoder as te
        from src.config import Config
        from src.multimodal_classifier import MultimodalClassifier

        from mrd_b200 import BIOBERT_BASE

        te.AutoConfig.from_pretrained = staticmethod(lambda name, **k: BertConfig(**BIOBERT_BASE))
        te.AutoModel.from_pretrained = staticmethod(lambda name, **k: BertModel(BertConfig(**BIOBERT_BASE)))
        cfg = Config()
        cfg.cnn_encoder.pretrained = False
        return MultimodalClassifier(cfg).eval()
    finally:
        os.chdir(cwd)


CASES = {
    # name: (weights, B, S, lengths, seed)
    "cfg1_plain_b4_s128": ("plain", 4, 128, None, 11),
    "cfg1_sens_b4_s128": ("sens", 4, 128, None, 12),
    "padded_sens_b5_s128": ("sens", 5, 128, [128, 77, 64, 17, 1], 13),
    "padded_sens_b3_s48": ("sens", 3, 48, [48, 33, 5], 14),
}
TEXT_CASES = {"text_sens_b2_s512": ("sens", 2, 512, [512, 300], 21),
              "text_plain_b3_s200": ("plain", 3, 200, [200, 129, 64], 22)}
IMAGE_CASES = {"image_sens_b2_160x96": ("sens", 2, 160, 96, 31)}


def _save(fix, path):
    torch.save({k: (v.detach().clone().contiguous() if torch.is_tensor(v) else v) for k, v in fix.items()},
               path)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    mine = synth.build_model(0)
    weights = {"plain": mine.state_dict()}
    weights["sens"] = synth.sensitise(weights["plain"], 1)
    ref = build_reference()
    meta = {"checksum": {k: synth.checksum(v) for k, v in weights.items()},
            "torch": str(torch.__version__)}
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        for name, (w, B, S, lengths, seed) in CASES.items():
            ref.load_state_dict(weights[w], strict=True)
            images, ids, mask = synth.make_inputs(B, S, seed, lengths)
            o = ref(images, ids, mask, return_embeddings=True)
            fix = {"weights": w, "B": B, "S": S, "lengths": lengths, "seed": seed,
                   "logits": o["logits"], "probs": o["probs"], "image_embedding": o["image_embedding"],
                   "text_embedding": o["text_embedding"], "fused_embedding": o["fused_embedding"],
                   "attn_i2t": o["attention_info"]["image_to_text_attention"],
                   "attn_t2i": o["attention_info"]["text_to_image_attention"]}
            _save(fix, os.path.join(out_dir, name + ".pt"))
            print(name, "logits std over samples", o["logits"].std(0).mean().item(),
                  "argmax", o["logits"].argmax(-1).tolist())
        for name, (w, B, S, lengths, seed) in TEXT_CASES.items():
            ref.load_state_dict(weights[w], strict=True)
            _, ids, mask = synth.make_inputs(B, S, seed, lengths, H=32, W=32)
            emb = ref.text_encoder(ids, mask)
            _save({"weights": w, "B": B, "S": S, "lengths": lengths, "seed": seed,
                   "text_embedding": emb}, os.path.join(out_dir, name + ".pt"))
            print(name, emb.std().item())
        for name, (w, B, H, W, seed) in IMAGE_CASES.items():
            ref.load_state_dict(weights[w], strict=True)
            images, _, _ = synth.make_inputs(B, 8, seed, None, H=H, W=W)
            fmap, emb = ref.cnn_encoder.get_intermediate_features(images)
            _save({"weights": w, "B": B, "H": H, "W": W, "seed": seed, "image_embedding": emb,
                   "pooled": fmap.mean(dim=(2, 3))}, os.path.join(out_dir, name + ".pt"))
            print(name, emb.std().item(), fmap.std().item())
    torch.save(meta, os.path.join(out_dir, "meta.pt"))
    print("wrote", sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main()
