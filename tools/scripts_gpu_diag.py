import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
import synth, mrd_b200
from oracle import forward_oracle as oracle
GOLD = "tests/golden"
model = synth.build_model(0)
plain = {k: v.clone() for k, v in model.state_dict().items()}
sens = synth.sensitise(plain, 1)
model = model.cuda()
W = {"plain": plain, "sens": sens}
def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-12)).max().item()
for name in ["cfg1_plain_b4_s128", "cfg1_sens_b4_s128", "padded_sens_b5_s128", "padded_sens_b3_s48"]:
    fix = torch.load(os.path.join(GOLD, name + ".pt"))
    model.load_state_dict(W[fix["weights"]])
    images, ids, mask = synth.make_inputs(fix["B"], fix["S"], fix["seed"], fix["lengths"])
    with torch.no_grad():
        out = model(images.cuda(), ids.cuda(), mask.cuda(), return_embeddings=True)
    lg = out["logits"].cpu()
    print(name, "img", rel(out["image_embedding"], fix["image_embedding"]), "txt", rel(out["text_embedding"], fix["text_embedding"]),
          "fused", rel(out["fused_embedding"], fix["fused_embedding"]), "logit abs", (lg - fix["logits"]).abs().max().item(),
          "logit rel", rel(lg, fix["logits"]), "logit max", fix["logits"].abs().max().item(),
          "fused norm", fix["fused_embedding"].norm(dim=-1).mean().item(), "top1", torch.equal(lg.argmax(-1), fix["logits"].argmax(-1)))
    # fusion+head alone on exact fp32 embeddings
    with torch.no_grad():
        fused, _ = model.fusion(fix["image_embedding"].cuda(), fix["text_embedding"].cuda())
        logits = model.classifier(fix["fused_embedding"].cuda())
    print("   fusion-only rel", rel(fused, fix["fused_embedding"]), "head-only abs", (logits.cpu() - fix["logits"]).abs().max().item())
