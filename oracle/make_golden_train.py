"""TEST INFRASTRUCTURE — generate tests/golden/istvt_golden_train.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden_train.py

One training iteration exactly as the ISTVT branch of train_CNN.py does it (lines 146-148,196-201,226,513-533):
the reference `XceptionVidTr` in train mode (BatchNorm batch statistics), `nn.BCEWithLogitsLoss`, `loss.backward()`,
`optim.AdamW(filter(requires_grad), lr, betas=(0.9, 0.999), eps=1e-8, weight_decay).step()`.
Stored: loss, logits, a fingerprint (shape, sum, max|x|, 256 sampled values) of the gradient of every parameter that
received one, of every such parameter after the optimizer step, and of the updated BatchNorm running statistics.
Weights are rebuilt from the seed (sensitised, see oracle/make_golden.py), never stored.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import istvt_oracle as O  # noqa: E402
from oracle import reference_shim  # noqa: E402
from oracle.make_golden import make_input  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "istvt_golden_train.pt")
SAMPLES = 256
LR, WD = 1e-3, 0.01
LABELS = [1, 0]


def fp(t: torch.Tensor) -> dict:
    t = t.detach().float().contiguous()
    flat = t.reshape(-1)
    idx = O.fingerprint_indices(flat.numel(), count=SAMPLES)
    return {"shape": tuple(t.shape), "sum": float(flat.double().sum()), "absmax": float(flat.abs().max()),
            "samples": flat[idx].clone()}


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 8)
    ref = reference_shim.build_reference_model(seed=0)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    O.sensitise_(sd)
    ref.load_state_dict(sd)
    ref.train()                                                           # train_CNN.py:226
    x = make_input(2, 6)
    labels = torch.tensor(LABELS)
    criterion = torch.nn.BCEWithLogitsLoss()                              # train_CNN.py:148
    params = filter(lambda p: p.requires_grad, ref.parameters())          # train_CNN.py:196
    opt = torch.optim.AdamW(params, lr=LR, betas=(0.9, 0.999), eps=1e-08, weight_decay=WD)   # train_CNN.py:199
    opt.zero_grad()
    out = ref(x)
    loss = criterion(out.view(-1), labels.float())                        # train_CNN.py:526
    loss.backward()
    named = dict(ref.named_parameters())
    grads = {k: fp(p.grad) for k, p in named.items() if p.grad is not None}
    no_grad = sorted(k for k, p in named.items() if p.grad is None)
    opt.step()
    sd_after = ref.state_dict()
    golden = {
        "torch_version": str(torch.__version__), "seed": 0, "sensitised": True, "batch": 2, "frames": 6, "labels": LABELS,
        "lr": LR, "weight_decay": WD, "loss": float(loss), "logits": out.detach().clone(),
        "grads": grads, "params_without_grad": no_grad,
        "params_after": {k: fp(sd_after[k]) for k in grads},
        "running_after": {k: fp(v) for k, v in sd_after.items()
                          if k.endswith(("running_mean", "running_var")) and k.split(".running")[0] + ".weight" in grads},
    }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(golden, OUT)
    print("loss", float(loss), "logits", out.flatten().tolist(), "grads", len(grads), "no-grad", len(no_grad))
    print("wrote", OUT, os.path.getsize(OUT) / 1e3, "KB")


if __name__ == "__main__":
    main()
