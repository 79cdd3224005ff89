import sys; sys.path.insert(0,'.')
import torch, ml4ca_b200 as M
dev=torch.device('cuda',0)
g=torch.Generator(device=dev); g.manual_seed(0)
for m in (1<<20, 1<<24, (1<<24)+1):
    eta=(torch.rand(3,m,device=dev,generator=g)*2-1)*torch.tensor([[8.0],[8.0],[0.785]],device=dev)
    nu=(torch.rand(3,m,device=dev,generator=g)*2-1)*torch.tensor([[1.4],[0.3],[0.52]],device=dev)
    ref,integ=torch.zeros(3,m,device=dev),torch.zeros(3,m,device=dev)
    for _ in range(3): M.pinv_pid(eta,nu,ref,integ)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): M.pinv_pid(eta,nu,ref,integ)
    e1.record(); torch.cuda.synchronize()
    t=e0.elapsed_time(e1)/20
    print("n %d: %.4f ms, %.0f GB/s of 80 B" % (m, t, 80*m/t/1e6))
