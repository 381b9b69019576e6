"""Pins oracle/forward_oracle.py (the CPU restatement) against the golden fixtures produced by the
UNMODIFIED reference (oracle/make_golden.py, run in the build container).  fp32 vs fp32: the only
differences are ATen kernel-selection / accumulation-order effects."""

import glob
import os

import pytest
import torch

import synth
from oracle import forward_oracle as oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def weights():
    plain = synth.build_model(0).state_dict()
    sens = synth.sensitise(plain, 1)
    meta = torch.load(os.path.join(GOLD, "meta.pt"))
    # the seeded construction must reproduce the exact tensors the fixtures were generated with
    assert synth.checksum(plain) == meta["checksum"]["plain"], "seeded weights differ from the fixture run"
    assert synth.checksum(sens) == meta["checksum"]["sens"]
    return {"plain": plain, "sens": sens}


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("name", ["cfg1_plain_b4_s128", "cfg1_sens_b4_s128", "padded_sens_b5_s128",
                                  "padded_sens_b3_s48"])
def test_full_forward_matches_reference(weights, name):
    fix = torch.load(os.path.join(GOLD, name + ".pt"))
    images, ids, mask = synth.make_inputs(fix["B"], fix["S"], fix["seed"], fix["lengths"])
    out = oracle.multimodal_forward(weights[fix["weights"]], images, ids, mask)
    for k in ("image_embedding", "text_embedding", "fused_embedding"):
        assert _rel(out[k], fix[k]) < 2e-5, (k, _rel(out[k], fix[k]))
    assert (out["logits"] - fix["logits"]).abs().max() < 1e-4  # north-star fp32 tolerance
    assert (out["probs"] - fix["probs"]).abs().max() < 1e-5
    assert torch.equal(out["logits"].argmax(-1), fix["logits"].argmax(-1))
    # softmax over a single key: exactly one, shape [B,8,1,1]
    for k, f in (("image_to_text_attention", "attn_i2t"), ("text_to_image_attention", "attn_t2i")):
        w = out["attention_info"][k]
        assert w.shape == fix[f].shape == (fix["B"], 8, 1, 1)
        assert torch.equal(w, fix[f]) and torch.equal(w, torch.ones_like(w))


@pytest.mark.parametrize("name", ["text_sens_b2_s512", "text_plain_b3_s200"])
def test_text_encoder_matches_reference(weights, name):
    fix = torch.load(os.path.join(GOLD, name + ".pt"))
    _, ids, mask = synth.make_inputs(fix["B"], fix["S"], fix["seed"], fix["lengths"], H=32, W=32)
    sd = {k: v.float() for k, v in weights[fix["weights"]].items() if v.is_floating_point()}
    with torch.no_grad():
        emb = oracle.text_encoder(sd, ids, mask)
    assert _rel(emb, fix["text_embedding"]) < 2e-5


def test_cnn_encoder_matches_reference(weights):
    fix = torch.load(os.path.join(GOLD, "image_sens_b2_160x96.pt"))
    images, _, _ = synth.make_inputs(fix["B"], 8, fix["seed"], None, H=fix["H"], W=fix["W"])
    sd = {k: v.float() for k, v in weights["sens"].items() if v.is_floating_point()}
    with torch.no_grad():
        pooled = oracle.resnet50_feature_map(sd, images).mean(dim=(2, 3))
        emb = oracle.cnn_encoder(sd, images)
    assert _rel(pooled, fix["pooled"]) < 2e-5
    assert _rel(emb, fix["image_embedding"]) < 2e-5


def test_oracle_invariances(weights):
    """Properties measured on the reference that the CUDA path must share."""
    sd = {k: v.float() for k, v in weights["sens"].items() if v.is_floating_point()}
    _, ids, mask = synth.make_inputs(2, 32, 5, [32, 9], H=32, W=32)
    with torch.no_grad():
        a = oracle.text_encoder(sd, ids, mask)
        ids2 = ids.clone()
        ids2[1, 9:] = 1234  # token ids at padded positions must not matter
        b = oracle.text_encoder(sd, ids2, mask)
        ones = torch.ones_like(mask)
        c = oracle.text_encoder(sd, ids, ones)
        d = oracle.text_encoder(sd, ids, None)
    assert torch.equal(a, b)
    assert torch.allclose(c, d, atol=1e-6)
    assert not torch.allclose(a[1], c[1], atol=1e-3)  # the mask does matter for the padded sample


def test_every_fixture_is_covered():
    names = {os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "*.pt"))} - {"meta"}
    assert names == {"cfg1_plain_b4_s128", "cfg1_sens_b4_s128", "padded_sens_b5_s128",
                     "padded_sens_b3_s48", "text_sens_b2_s512", "text_plain_b3_s200",
                     "image_sens_b2_160x96",
                     # training-step fixtures: covered by tests/test_train_oracle.py
                     "train_p0_bn_eval_b4_s32", "train_p0_bn_train_b4_s32"}
