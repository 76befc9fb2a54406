#!/usr/bin/env python
"""Long sequences on the small-batch regime: R3 against R2 on identical inputs (T = 128 and 300: many wraps of the activation
ring, the accumulator double buffer and the group-barrier epochs).  Development aid.  usage: long_seq_r3_vs_r2.py"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from vmlmf_b200 import _lib
    from vmlmf_b200.functional import vmlmf_sequence
    out = {}
    for (T, B, I, H, RX, RH) in [(128, 32, 9, 1024, 64, 64), (300, 20, 24, 650, 40, 300), (257, 5, 8, 200, 20, 130)]:
        g = torch.Generator(device="cuda").manual_seed(T)
        r = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.05).requires_grad_(True)
        canon = [r(I, RX), r(4 * H, RX), r(4, I), r(H, RH), r(4 * H, RH), r(4, H), r(4 * H)]
        x = torch.randn(T, B, I, device="cuda", generator=g)
        dy = torch.randn(T, B, H, device="cuda", generator=g)
        y, hT, cT = vmlmf_sequence(x, None, None, canon, False)
        torch.autograd.backward([y, cT], [dy, torch.ones_like(cT)])
        torch.cuda.synchronize()
        key = f"{T},{B},{I},{H},{RX},{RH}"
        out[key] = {"path": _lib.plan(T, B, I, H, RX, RH).path, "y": y.double().norm().item(), "ysum": y.double().sum().item(),
                    "cT": cT.double().sum().item(), "grads": [p.grad.double().norm().item() for p in canon],
                    "finite": bool(torch.isfinite(y).all() and all(torch.isfinite(p.grad).all() for p in canon))}
        torch.save([y.cpu(), cT.cpu()] + [p.grad.cpu() for p in canon], f"/tmp/lsr_{os.environ.get('VMLMF_NO_R3','0')}_{T}.pt")
    print(json.dumps(out))
else:
    import torch
    res = []
    for env in ({}, {"VMLMF_NO_R3": "1"}):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "run"], capture_output=True, text=True, env=e)
        if r.returncode:
            print(r.stderr[-1500:]); sys.exit(1)
        res.append(json.loads(r.stdout.strip().splitlines()[-1]))
    worst = 0.0
    for key in res[0]:
        T = key.split(",")[0]
        a = torch.load(f"/tmp/lsr_0_{T}.pt"); b = torch.load(f"/tmp/lsr_1_{T}.pt")
        errs = [((u.double() - v.double()).norm() / v.double().norm().clamp_min(1e-30)).item() for u, v in zip(a, b)]
        worst = max(worst, max(errs))
        print(key, "paths", res[0][key]["path"], res[1][key]["path"], "finite", res[0][key]["finite"], res[1][key]["finite"],
              "max rel-l2 R3 vs R2 over y, cT, 7 grads:", f"{max(errs):.2e}")
    sys.exit(0 if worst < 1e-5 else 1)
