"""Pin oracle/torch_port.py (the CPU arm bench.py times) to the reference run (golden vectors)."""
import numpy as np
import torch

from oracle.torch_port import PortCriterionStep


def test_port_matches_reference_run(golden):
    g = golden("criterion_step")
    torch.set_num_threads(1)
    port = PortCriterionStep(s_dim=24, t_dim=24, feat_dim=32, K=40, T=0.15, alpha=0.999, num_heads=4,
                             ema_shapes=[(3, 5)])
    mods = {"embed_s": port.embed_s, "embed_t": port.embed_t, "atts_q": port.atts_q, "atts_k": port.atts_k,
            "atts_queue": port.atts_queue}

    def load(prefix):
        for name, mod in mods.items():
            sd = {k[len(name) + 1:]: torch.from_numpy(g[prefix + k].copy()) for k in
                  [f[len(prefix):] for f in g.files if f.startswith(prefix + name + ".")]}
            mod.load_state_dict(sd)

    load("sd0_")
    port.contrast.memory.copy_(torch.from_numpy(g["mem0"]))
    for st in range(3):
        fs = torch.from_numpy(g[f"st{st}_feat_s"].copy()).requires_grad_()
        ft = torch.from_numpy(g[f"st{st}_feat_t"].copy())
        loss, acc = port.step(fs, ft)
        assert abs(loss.item() - float(g[f"st{st}_loss"])) < 1e-6 * abs(float(g[f"st{st}_loss"]))
        assert acc.item() == float(g[f"st{st}_acc"][0])
        assert np.allclose(fs.grad.numpy(), g[f"st{st}_dfeat_s"], rtol=1e-5, atol=1e-8)
        assert np.array_equal(port.contrast.memory.numpy(), g[f"st{st}_mem"])
        assert port.contrast.index == int(g[f"st{st}_index"])
        assert port.atts_k.qkv.weight.grad is None and port.atts_queue.proj.weight.grad is None   # KAT6
        load(f"st{st}_sd_")          # continue from the reference's post-SGD parameters
