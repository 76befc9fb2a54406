set -x
( timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r2b_pytest.log
python - <<'PY'
import torch, statistics
import vmlmf_b200 as vb
from vmlmf_b200 import functional as F
dev="cuda:0"
for (I,H,B,T) in ((9,128,64,128),(77,256,81,24),(9,128,16,128)):
    torch.manual_seed(3)
    net=vb.Net(I,[H],w_rank=8,u_rank=[6],cell=vb.MyVMLMFCell).to(dev)
    x=torch.randn(B,T,I,device=dev); y=torch.randint(0,6,(B,),device=dev)
    for _ in range(5): vb.cross_entropy(net(x),y).backward()
    F.EVENT_LOG=[]
    for _ in range(20):
        net.zero_grad(); vb.cross_entropy(net(x),y).backward()
    torch.cuda.synchronize()
    lat={}
    for nm,a,b in F.EVENT_LOG: lat.setdefault(nm,[]).append(a.elapsed_time(b))
    F.EVENT_LOG=None
    print((I,H,B,T), "fwd us/step %.3f bwd us/step %.3f" % (statistics.median(lat["seq_fwd"])*1e3/T, statistics.median(lat["seq_bwd"])*1e3/T))
PY
