"""fp32 check mode (BASELINE.json north_star: "logits within ... 1e-4 in an fp32 check mode").

configure_b200(fp32_check=True) routes the same module API / C-ABI entry points through the library's
plain-fp32 CUDA kernels (csrc/fp32_check.cu, csrc/simt_gemm.cu: no bf16, BatchNorm un-folded).  The
comparator is the unmodified reference itself: tests/golden/*.pt were produced by running
/root/reference on CPU in fp32 (oracle/make_golden.py).  Tolerances, written here as the spec asks:
  logits / probs max-abs <= 1e-4 on random-init weights (measured ~1e-6);
  per-sample relative L2 <= 1e-4 on every embedding and on the sensitised logits (|logits| ~ 14, so a
  max-abs bar would be meaningless there); identical top-1.
"""

import os

import pytest
import torch

import synth
from oracle import forward_oracle as oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ABS_TOL = 1e-4
REL_TOL = 1e-4


def _rel_rows(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-12)).max().item()


@pytest.fixture(scope="module")
def checked():
    model = synth.build_model(0)
    plain = {k: v.clone() for k, v in model.state_dict().items()}
    sens = synth.sensitise(plain, 1)
    model = model.to("cuda:0")
    for m in (model, model.cnn_encoder, model.text_encoder, model.fusion, model.classifier):
        m.configure_b200(fp32_check=True)
    return {"model": model, "plain": plain, "sens": sens, "loaded": "plain"}


def _use(st, which):
    if st["loaded"] != which:
        st["model"].load_state_dict(st[which], strict=True)
        st["loaded"] = which
    return st["model"]


@pytest.mark.parametrize("name", ["cfg1_plain_b4_s128", "cfg1_sens_b4_s128", "padded_sens_b5_s128",
                                  "padded_sens_b3_s48"])
def test_fp32_check_full_forward_vs_reference(cuda, checked, name):
    fix = torch.load(os.path.join(GOLD, name + ".pt"))
    model = _use(checked, fix["weights"])
    images, ids, mask = synth.make_inputs(fix["B"], fix["S"], fix["seed"], fix["lengths"])
    with torch.no_grad():
        out = model(images.cuda(), ids.cuda(), mask.cuda(), return_embeddings=True)
    torch.cuda.synchronize()
    errs = {k: _rel_rows(out[k], fix[k]) for k in ("image_embedding", "text_embedding", "fused_embedding",
                                                   "logits")}
    print(name, {k: f"{v:.2e}" for k, v in errs.items()},
          "logits max-abs %.2e" % (out["logits"].cpu() - fix["logits"]).abs().max().item())
    for k, v in errs.items():
        assert v <= REL_TOL, (k, v)
    if fix["weights"] == "plain":
        assert (out["logits"].cpu() - fix["logits"]).abs().max().item() <= ABS_TOL
    assert (out["probs"].cpu() - fix["probs"]).abs().max().item() <= ABS_TOL
    assert torch.equal(out["logits"].argmax(-1).cpu(), fix["logits"].argmax(-1))
    w = out["attention_info"]["image_to_text_attention"].cpu()
    assert torch.equal(w, fix["attn_i2t"])


def test_fp32_check_text_encoder_s512(cuda, checked):
    fix = torch.load(os.path.join(GOLD, "text_sens_b2_s512.pt"))
    model = _use(checked, fix["weights"])
    _, ids, mask = synth.make_inputs(fix["B"], fix["S"], fix["seed"], fix["lengths"], H=32, W=32)
    with torch.no_grad():
        emb = model.text_encoder(ids.cuda(), mask.cuda())
    assert _rel_rows(emb, fix["text_embedding"]) <= REL_TOL


def test_fp32_check_cnn_encoder_odd_size(cuda, checked):
    fix = torch.load(os.path.join(GOLD, "image_sens_b2_160x96.pt"))
    model = _use(checked, "sens")
    images, _, _ = synth.make_inputs(fix["B"], 8, fix["seed"], None, H=fix["H"], W=fix["W"])
    with torch.no_grad():
        fmap, emb = model.cnn_encoder.get_intermediate_features(images.cuda())
    assert _rel_rows(fmap.mean(dim=(2, 3)), fix["pooled"]) <= REL_TOL
    assert _rel_rows(emb, fix["image_embedding"]) <= REL_TOL


def test_fp32_check_brackets_the_bf16_path(cuda, checked):
    """The check mode and the fast path see the same weights and inputs: their difference is the bf16
    rounding of the fast path alone (no structural disagreement), far inside the 2e-2 bar."""
    model = _use(checked, "sens")
    images, ids, mask = synth.make_inputs(6, 64, 31, [64, 9, 33, 1, 50, 17])
    args = (images.cuda(), ids.cuda(), mask.cuda())
    with torch.no_grad():
        ref = model(*args, return_embeddings=True)
        model.configure_b200(fp32_check=False)
        try:
            fast = model(*args, return_embeddings=True)
        finally:
            model.configure_b200(fp32_check=True)
    sd = {k: v.float() for k, v in checked["sens"].items() if v.is_floating_point()}
    want = oracle.multimodal_forward(sd, images, ids, mask)
    for k in ("image_embedding", "text_embedding", "fused_embedding", "logits"):
        assert _rel_rows(ref[k], want[k]) <= REL_TOL, (k, _rel_rows(ref[k], want[k]))
        if k == "logits":   # the bf16 bar, relative to the logit scale of the sensitised weights (|x| ~ 14)
            err = (fast[k] - ref[k]).abs().max().item()
            assert err <= 2e-2 * max(1.0, ref[k].abs().max().item()), err
        else:
            assert _rel_rows(fast[k], ref[k]) <= 2e-2, (k, _rel_rows(fast[k], ref[k]))
        assert _rel_rows(fast[k], ref[k]) > 1e-6   # they really are two different arithmetic paths
