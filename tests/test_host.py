"""Host-side logic that needs no GPU: drop-in API surface, state_dict contract, loud failure without
CUDA, shard arithmetic and the world_size-2 logits gather (gloo)."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mrd_b200
import synth


@pytest.fixture(scope="module")
def model():
    return synth.build_model(0)


def test_state_dict_contract(model):
    sd = model.state_dict()
    assert len(sd) == 555  # SURVEY.md section 5 (PROBE on the reference)
    assert sum(p.numel() for p in model.parameters()) == 136_842_698
    for k in ("cnn_encoder.backbone.conv1.weight", "cnn_encoder.backbone.layer4.2.bn3.running_var",
              "cnn_encoder.projection.0.weight", "cnn_encoder.projection.3.bias",
              "text_encoder.encoder.embeddings.word_embeddings.weight",
              "text_encoder.encoder.encoder.layer.11.output.LayerNorm.bias",
              "text_encoder.encoder.pooler.dense.weight",
              "fusion.fusion_layer.image_to_text_attention.query_proj.weight",
              "fusion.fusion_layer.text_to_image_attention.output_proj.bias",
              "fusion.fusion_layer.layer_norm_text.weight", "fusion.fusion_layer.fusion.3.weight",
              "classifier.classifier.0.weight", "classifier.classifier.3.weight",
              "classifier.classifier.6.bias"):
        assert k in sd, k
    assert sd["classifier.classifier.6.weight"].shape == (10, 128)
    assert sd["fusion.fusion_layer.text_proj.weight"].shape == (512, 768)


def test_module_tree_matches_reference_walks(model):
    # the trainers walk these (src/train_multimodal.py:429-493); notebooks hook backbone.layer4
    assert [n for n, _ in model.cnn_encoder.backbone.named_children()] == [
        "conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4", "avgpool", "fc"]
    assert len(model.text_encoder.encoder.encoder.layer) == 12
    assert model.cnn_encoder.get_attention_layer() is model.cnn_encoder.backbone.layer4
    assert model.num_classes == 10 and model.fusion_dim == 512
    assert model.image_embedding_dim == 512 and model.text_embedding_dim == 768
    # default config freezes the backbone (src/config.py:64) but not the projection
    assert not any(p.requires_grad for p in model.cnn_encoder.backbone.parameters())
    assert all(p.requires_grad for p in model.cnn_encoder.projection.parameters())
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert trainable == 136_842_698 - 23_508_032


def test_no_cpu_fallback(model):
    x = torch.zeros(1, 3, 224, 224)
    ids = torch.zeros(1, 8, dtype=torch.long)
    with pytest.raises(mrd_b200.MrdError, match="CUDA devices only"):
        model(x, ids, torch.ones_like(ids))
    with pytest.raises(mrd_b200.MrdError):
        model.cnn_encoder(x)
    with pytest.raises(mrd_b200.MrdError):
        model.text_encoder(ids, torch.ones_like(ids))
    with pytest.raises(mrd_b200.MrdError):
        model.fusion(torch.zeros(1, 512), torch.zeros(1, 768))
    with pytest.raises(mrd_b200.MrdError):
        model.classifier(torch.zeros(1, 512))


def test_training_mode_never_falls_back(model):
    """Train mode: MultimodalClassifier.forward is the library's training step (CUDA only - on a CPU model it
    fails like every other forward); the standalone sub-modules have no train-mode path and say so."""
    model.train()
    try:
        with pytest.raises(mrd_b200.MrdError, match="CUDA"):
            model(torch.zeros(1, 3, 224, 224), torch.zeros(1, 8, dtype=torch.long),
                  torch.ones(1, 8, dtype=torch.long))
        with pytest.raises(NotImplementedError, match="eval"):
            model.cnn_encoder(torch.zeros(1, 3, 224, 224))
        # return_embeddings is served by the training step too (same CUDA-only rule)
        with pytest.raises(mrd_b200.MrdError, match="CUDA"):
            model(torch.zeros(1, 3, 224, 224), torch.zeros(1, 8, dtype=torch.long),
                  torch.ones(1, 8, dtype=torch.long), return_embeddings=True)
    finally:
        model.eval()


def test_unsupported_configs_are_rejected():
    with pytest.raises(ValueError, match="Unknown backbone"):
        mrd_b200.CNNEncoder(mrd_b200.CNNEncoderConfig(backbone="vgg", pretrained=False))
    with pytest.raises(NotImplementedError):
        mrd_b200.CNNEncoder(mrd_b200.CNNEncoderConfig(backbone="efficientnet_b0", pretrained=False))
    with pytest.raises(ValueError, match="Unknown fusion type"):
        mrd_b200.MultimodalFusion(mrd_b200.FusionConfig(fusion_type="bilinear"))
    with pytest.raises(NotImplementedError):
        mrd_b200.MultimodalFusion(mrd_b200.FusionConfig(fusion_type="gated"))


def test_freeze_helpers():
    enc = mrd_b200.CNNEncoder(mrd_b200.CNNEncoderConfig(pretrained=False, freeze_backbone=False,
                                                        freeze_layers=2))
    frozen = {n.split(".")[0] for n, p in enc.backbone.named_parameters() if not p.requires_grad}
    assert frozen == {"conv1", "bn1", "layer1", "layer2"}


def test_deepcopy_and_pickle_drop_engine(model):
    import copy
    import pickle

    m2 = copy.deepcopy(model.classifier)
    assert "_mrd_engine" not in m2.__dict__
    pickle.loads(pickle.dumps(model.classifier))


def test_getstate_drops_runtime_state(model):
    """Pickles / deep copies carry parameters only: no engine, copy stream, prefetched device tensors or process
    group, and no cache of applied train options (that lives on the Engine, so a new engine gets them again)."""
    import copy

    model.__dict__["_mrd_copy_stream"] = object()
    model.__dict__["_mrd_prefetched"] = {"x": 1}
    model.__dict__["_mrd_ddp"] = (True, None)
    try:
        m2 = copy.deepcopy(model)
        for k in ("_mrd_engine", "_mrd_copy_stream", "_mrd_prefetched", "_mrd_ddp", "_mrd_train_opts"):
            assert k not in m2.__dict__, k
    finally:
        for k in ("_mrd_copy_stream", "_mrd_prefetched", "_mrd_ddp"):
            model.__dict__.pop(k, None)


def test_predict_batch_restores_per_module_modes(model, monkeypatch):
    """predict_batch_tensors puts back every module's own training flag (backbone.eval() inside model.train())."""
    from importlib import import_module
    pred = import_module("multimodal-rare-disease_b200.predict")
    model.train()
    model.cnn_encoder.backbone.eval()
    try:
        monkeypatch.setattr(type(model), "forward",
                            lambda self, images=None, input_ids=None, attention_mask=None, **k:
                            {"probs": torch.full((images.shape[0], 10), 0.1)})
        # a non-CPU, non-CUDA device tag routes through model(...) (the patched forward) instead of forward_host
        images = torch.zeros(2, 3, 8, 8, device="meta")
        pred.predict_batch_tensors(model, images, torch.zeros(2, 4, dtype=torch.long), None)
        assert model.training and model.text_encoder.training
        assert not model.cnn_encoder.backbone.training
        assert not any(m.training for m in model.cnn_encoder.backbone.modules())
    finally:
        model.eval()


def test_shard_bounds():
    for total in (0, 1, 7, 16, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [mrd_b200.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mrd_b200.shard_bounds(4, 2, 2)


def _dp_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        images = torch.randn(total, 3, 4, 4, generator=g)
        ids = torch.randint(0, 100, (total, 6), generator=g)
        mask = torch.ones(total, 6, dtype=torch.long)

        def local_forward(x, i, m, logits_out):  # stand-in for the CUDA forward: a per-sample function
            logits_out.copy_(x.flatten(1)[:, :10] + i[:, :1].float() + m.sum(1, keepdim=True))

        dp = mrd_b200.DataParallelForward(local_forward, 10)
        got = dp.forward_global(images, ids, mask)
        want = images.flatten(1)[:, :10] + ids[:, :1].float() + 6.0
        lo, hi = mrd_b200.shard_bounds(total, world, rank)
        got2 = dp.forward_shard(images[lo:hi], ids[lo:hi], mask[lo:hi], total)
        q.put((rank, torch.equal(got, want), torch.equal(got2, want), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_data_parallel_gather_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok1, ok2, shape in res:
        assert ok1 and ok2 and shape == (total, 10), (rank, ok1, ok2, shape)


def _ar_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from importlib import import_module
        par = import_module("multimodal-rare-disease_b200.parallel")
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
        views = [flat[:4].view(2, 2), flat[4:]]          # parameter gradients are views of the bucket
        mrd_b200.allreduce_mean_(flat)
        want = torch.arange(10, dtype=torch.float32) * 1.5
        ok = torch.equal(flat, want) and torch.equal(views[0], want[:4].view(2, 2))
        # bucketed variant (the overlapped path): same result bucket by bucket
        flat2 = torch.arange(10, dtype=torch.float32) * (rank + 1)
        works = [par.allreduce_mean_async(flat2[a:b]) for a, b in ((6, 10), (2, 6), (0, 2))]
        for w in works:
            par.wait_allreduce(w)
        ok = ok and torch.equal(flat2, want)
        # per-rank dropout seeds differ, rank 0 keeps the caller's seed
        s = par.rank_seed(12345)
        ok = ok and ((s == 12345) if rank == 0 else (s != 12345)) and 0 <= s < 2 ** 62
        # BatchNorm buffers: rank 0's values everywhere
        bufs = [torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1)))]
        par.broadcast_buffers_(bufs)
        ok = ok and torch.equal(bufs[0], torch.ones(3)) and torch.equal(bufs[1], torch.full((2, 2), 10.0))
        q.put((rank, ok, True))
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_allreduce_world2():
    """The training step's only collective: mean of the flat gradient bucket over the ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ar_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)
    assert mrd_b200.allreduce_mean_(torch.ones(3)) is not None   # no process group: a no-op


def test_fused_adamw_host_contract():
    """FusedAdamW is a torch.optim.Optimizer with AdamW's param-group / state_dict layout; it has no CPU path."""
    lin = torch.nn.Linear(4, 3)
    groups = [{"params": [lin.weight], "lr": 1e-3}, {"params": [lin.bias], "lr": 1e-2, "weight_decay": 0.0}]
    opt = mrd_b200.FusedAdamW(groups, lr=5e-5, weight_decay=0.05, max_grad_norm=1.0)
    assert isinstance(opt, torch.optim.Optimizer)
    assert opt.param_groups[0]["weight_decay"] == 0.05 and opt.param_groups[1]["weight_decay"] == 0.0
    ref = torch.optim.AdamW([{"params": [lin.weight], "lr": 1e-3}, {"params": [lin.bias], "lr": 1e-2, "weight_decay": 0.0}],
                            lr=5e-5, weight_decay=0.05)
    lin(torch.randn(2, 4)).sum().backward()
    ref.step()
    opt.load_state_dict(ref.state_dict())          # a torch AdamW checkpoint loads
    assert set(opt.state_dict()["state"][0]) >= {"step", "exp_avg", "exp_avg_sq"}
    with pytest.raises(mrd_b200.MrdError, match="CUDA"):
        opt.step()
    with pytest.raises(ValueError):
        mrd_b200.FusedAdamW([lin.weight], lr=-1.0)


def test_predict_batch_formatting_matches_reference_semantics():
    """format_predictions == the per-sample dicts of MultimodalPredictor.predict_batch (src/predict.py:241-267),
    including numpy's argsort()[::-1] tie order; predict_batch_tensors drives a model object through it."""
    import numpy as np

    probs = torch.tensor([[0.1, 0.5, 0.2, 0.2], [0.25, 0.25, 0.25, 0.25], [0.7, 0.1, 0.1, 0.1]])
    names = ["A", "B", "C"]          # shorter than the class count: the reference falls back to Class_i
    got = mrd_b200.format_predictions(probs, names, top_k=3)
    for i, sample in enumerate(probs.numpy()):
        top = sample.argsort()[::-1][:3]
        want = [{"syndrome": names[j] if j < len(names) else f"Class_{j}", "class_id": int(j),
                 "confidence": float(sample[j])} for j in top]
        assert got[i]["sample_idx"] == i and got[i]["predictions"] == want and got[i]["top_prediction"] == want[0]
    assert mrd_b200.format_predictions(probs, None, top_k=0)[0]["top_prediction"] is None

    class Fake:
        training = True
        calls = []

        def eval(self):
            self.training = False

        def train(self, mode=True):
            self.training = mode

        def modules(self):
            return [self]

        def forward_host(self, images, ids, mask, micro_batch=512):
            self.calls.append(("host", micro_batch, self.training))
            return {"probs": probs[: images.shape[0]]}

        def __call__(self, images, input_ids, attention_mask):
            self.calls.append(("dev", self.training))
            return {"probs": probs[: images.shape[0]]}

    m = Fake()
    res = mrd_b200.predict_batch_tensors(m, torch.zeros(2, 3, 8, 8), torch.zeros(2, 4, dtype=torch.long), None,
                                         names, top_k=1, micro_batch=64)
    assert m.calls == [("host", 64, False)] and m.training is True and len(res) == 2
    assert res[1]["top_prediction"]["class_id"] == int(np.argsort(probs[1].numpy())[::-1][0])
    with pytest.raises(ValueError, match="must match"):
        mrd_b200.predict_batch_tensors(m, torch.zeros(2, 3, 8, 8), torch.zeros(3, 4, dtype=torch.long), None)


def test_weight_handover_groups(model):
    """Parameters are handed to the library in five independent groups (a training step re-packs the trainable
    ones only); every floating-point state_dict entry lands in exactly one group."""
    from importlib import import_module

    Engine = import_module("multimodal-rare-disease_b200.engine").Engine
    named = list(model._mrd_named())
    groups = Engine.group_named(named)
    assert list(groups) == ["cnn_encoder.backbone.", "cnn_encoder.projection.", "text_encoder.", "fusion.", "classifier."]
    n_float = sum(1 for _, t in named if t.is_floating_point())
    assert sum(len(v) for v in groups.values()) == n_float
    assert len(groups["cnn_encoder.projection."]) == 4 and len(groups["classifier."]) == 6
    assert all(n.startswith("cnn_encoder.backbone.") for n, _ in groups["cnn_encoder.backbone."])
    # the frozen backbone is exactly the group a default-configuration training step never re-packs
    assert not any(t.requires_grad for _, t in groups["cnn_encoder.backbone."])
    assert all(t.requires_grad for _, t in groups["cnn_encoder.projection."])


def test_collect_predictions_and_checkpoint_loader(tmp_path):
    """collect_predictions == Evaluator.collect_predictions (src/evaluate.py:79-123) on any model with the reference's
    output dict; load_checkpoint == MultimodalPredictor._load_checkpoint (src/predict.py:73-82)."""
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(3 * 4 * 4, 5)

        def forward(self, images, input_ids=None, attention_mask=None):
            logits = self.lin(images.flatten(1))
            if input_ids is not None:
                logits = logits + input_ids[:, :5].float() * attention_mask[:, :5].float()
            return {"logits": logits, "probs": torch.softmax(logits, -1)}

    torch.manual_seed(3)
    m = Tiny().train()
    batches = [{"image": torch.randn(4, 3, 4, 4), "input_ids": torch.randint(0, 9, (4, 6)),
                "attention_mask": torch.ones(4, 6, dtype=torch.long), "label": torch.tensor([0, 1, 2, 3])},
               {"image": torch.randn(2, 3, 4, 4), "input_ids": torch.randint(0, 9, (2, 6)),
                "attention_mask": torch.ones(2, 6, dtype=torch.long), "label": [4, 0]}]
    preds, labels, probs = mrd_b200.collect_predictions(m, batches, mode="multimodal", device="cpu")
    assert m.training                                    # the caller's mode flags are restored
    assert preds.shape == (6,) and labels.tolist() == [0, 1, 2, 3, 4, 0] and probs.shape == (6, 5)
    with torch.no_grad():
        want = torch.cat([m.eval()(b["image"], b["input_ids"], b["attention_mask"])["probs"] for b in batches])
    assert torch.allclose(probs, want) and torch.equal(preds, want.argmax(-1))
    p2, l2, _ = mrd_b200.collect_predictions(m, [(b["image"], b["label"]) for b in batches[:1]], mode="image_only",
                                             device="cpu")
    assert p2.shape == (4,) and l2.tolist() == [0, 1, 2, 3]
    with pytest.raises(ValueError, match="Unknown mode"):
        mrd_b200.collect_predictions(m, batches, mode="audio")
    # checkpoints as src/train.py:394-437 writes them, and bare state_dicts
    path = tmp_path / "best_model.pt"
    torch.save({"model_state_dict": m.state_dict(), "epoch": 7}, path)
    fresh = Tiny()
    ck = mrd_b200.load_checkpoint(fresh, path)
    assert ck["epoch"] == 7 and torch.equal(fresh.lin.weight, m.lin.weight)
    torch.save(m.state_dict(), tmp_path / "bare.pt")
    mrd_b200.load_checkpoint(Tiny(), tmp_path / "bare.pt")
    with pytest.raises(FileNotFoundError, match="Checkpoint not found"):
        mrd_b200.load_checkpoint(fresh, tmp_path / "missing.pt")


def test_pair_tile_schedule_covers_every_tile_once():
    """The persistent tile schedule of the cta_group::2 kernels (csrc/gemm_conv.cu, `tile_at`), restated: cluster c of P
    takes pair-tiles j = c + it * P of ceil(stripes / 2) x n_tiles; j = (stripe pair, column tile); the CTA of rank r
    works on stripe 2 * pair + r.  Every real tile must be produced exactly once, both CTAs of a cluster must run the
    same number of iterations on the same column tile (they share the weight block), and a tile outside the matrix
    appears only as the second stripe of the last pair when the stripe count is odd."""
    for stripes, ntn, P in ((579, 9, 74), (579, 3, 74), (8, 12, 74), (1, 3, 74), (150, 1, 74), (455, 1, 74), (3, 2, 5)):
        total = stripes * ntn
        pair_total = (stripes + 1) // 2 * ntn
        seen = {}
        for c in range(P):
            my_tiles = (pair_total - c + P - 1) // P if pair_total > c else 0
            for it in range(my_tiles):
                j = c + it * P
                pm, n = divmod(j, ntn)
                tiles = [(2 * pm + r) * ntn + n for r in (0, 1)]
                assert tiles[0] % ntn == tiles[1] % ntn == n
                assert tiles[0] < total, "the leader's tile is always inside the matrix"
                for t in tiles:
                    if t < total:
                        seen[t] = seen.get(t, 0) + 1
                    else:
                        assert stripes % 2 == 1 and t // ntn == stripes, "only the odd stripe count has a phantom stripe"
        assert sorted(seen) == list(range(total)) and set(seen.values()) == {1}, (stripes, ntn, P)
