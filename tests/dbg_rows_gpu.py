import os,sys,time
sys.path.insert(0,os.getcwd())
import numpy as np, torch
import complogic_b200 as host
from complogic_b200 import circuits
from oracle import oracle
from tests.helpers import mask_tail, random_sliced, stride_for
ctx=host.Context(0)
# small programs first
for name,circ,nv in (("dag64x1024",circuits.random_dag(64,1024,256,2,3),1<<16),("dag40x512",circuits.random_dag(40,512,128,2,5),1<<15)):
    imm=random_sliced(circ.n_immediates, stride_for(nv), seed=4)
    want=mask_tail(oracle.run_batch_sliced(circ.sim,len(circ.sim.registers),imm,nv,circ.out_regs),nv)
    cs=host.CudaSimulation.from_simulation(circ.sim,circ.n_immediates,circ.out_regs,ctx=ctx,mode=host.MODE_COOP)
    got=mask_tail(cs.run_batch(imm,nv),nv)
    bad=np.argwhere(got!=want)
    print(name, cs.info.get("rows_segments"), "OK" if len(bad)==0 else f"MISMATCH {len(bad)}", flush=True)
circ=circuits.random_dag()
cs=host.CudaSimulation.from_simulation(circ.sim,1024,circ.out_regs,ctx=ctx,mode=host.MODE_COOP)
for lg in (18,20,22,22):
    nv, period = 1<<lg, 1<<12
    stride, pw = stride_for(nv), period//32
    block=random_sliced(1024,pw,seed=0xD4A)
    imm=torch.from_numpy(block.view(np.int32)).cuda().repeat(1,stride//pw).contiguous()
    out=torch.zeros((4096,stride),dtype=torch.int32,device="cuda:0")
    cs.run_device(imm.data_ptr(),out.data_ptr(),nv,stride)
    torch.cuda.synchronize()
    blocks=out.view(4096,stride//pw,pw)
    ne=(blocks!=blocks[:,:1,:])
    nbad=int(ne.sum().item())
    first=blocks[:,0,:].contiguous().cpu().numpy().view(np.uint32)
    want=oracle.run_batch_sliced(circ.sim,len(circ.sim.registers),block,period,circ.out_regs)
    print("lg",lg,"mismatching words vs first block",nbad,"first block ok",np.array_equal(first,want),flush=True)
    if nbad:
        idx=ne.nonzero()[:20].cpu().numpy()
        print(" sample (row, block, word):",idx.tolist())
        bb=ne.any(dim=0).any(dim=1).nonzero().flatten().cpu().numpy()
        print(" bad blocks:",len(bb),bb[:40].tolist())
        rr=ne.any(dim=2).any(dim=1).nonzero().flatten().cpu().numpy()
        print(" bad rows:",len(rr),rr[:40].tolist())
