import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch, vmlmf_b200 as vb
    torch.manual_seed(5)
    dev = "cuda:0"
    m = vb.Model(10000, 650, 2, 0.0, 0.05, w_rank=300, u_ranks=[300], lstm_type="vmlmf").to(dev)
    opt = vb.FlatClipSGD(m, lr=1.0, max_norm=5.0)
    g = torch.Generator().manual_seed(3)
    B, T = 20, 35
    states = None
    losses = []
    for it in range(12):
        tok = torch.randint(0, 10000, (T, B), generator=g).to(dev)
        y = torch.randint(0, 10000, (T, B), generator=g).to(dev)
        if states is None:
            states = [(torch.zeros(B, 650, device=dev), torch.zeros(B, 650, device=dev)) for _ in range(2)]
        states = m.detach(states)
        opt.zero_grad()
        s, states = m(tok, states)
        loss = vb.nll_loss(s, y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    from vmlmf_b200 import _lib
    print(json.dumps({"path": _lib.plan(T, B, 650, 650, 300, 300).path, "losses": losses,
                      "pnorm": float(torch.cat([p.detach().reshape(-1) for p in m.parameters()]).norm())}))
else:
    outs = []
    for env in ({}, {"VMLMF_NO_R3": "1"}):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "run"], capture_output=True, text=True, env=e)
        print(r.stderr[-500:] if r.returncode else "", end="")
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    a, b = outs
    print("paths", a["path"], b["path"])
    for i, (x, y) in enumerate(zip(a["losses"], b["losses"])):
        print(i, f"{x:.6f} {y:.6f} rel {abs(x-y)/abs(y):.2e}")
    print("param norm", a["pnorm"], b["pnorm"], abs(a["pnorm"] - b["pnorm"]) / b["pnorm"])
