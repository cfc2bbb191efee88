import sys, torch, time
sys.path.insert(0, ".")
from esmdiff_b200.engine import Dims, Engine
from esmdiff_b200.synthetic import random_state_dict
from esmdiff_b200.tokenization import synthetic_sequence_tokens
dev = torch.device("cuda")
eng = Engine(Dims()); eng.load_state_dict(random_state_dict(Dims(), device=dev, seed=0))
sched = eng.schedule(25)
for L, N in ((128, 64), (128, 512), (256, 512)):
    T = L + 2
    seq = synthetic_sequence_tokens(L, seed=0).to(dev)[None].expand(N, T).contiguous()
    eng.ddpm_sample(seq, None, 2, *eng.schedule(2), seed=1); torch.cuda.synchronize()
    for prof in (True, False):
        eng.profile(prof)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = eng.ddpm_sample(seq, None, 25, *sched, seed=2); e1.record(); torch.cuda.synchronize(); eng.synchronize()
        ms = e0.elapsed_time(e1)
        line = f"L={L} N={N} prof={prof}: {ms:.1f} ms {N*L/ms*1e3:.0f} tok/s"
        if prof:
            eng.profile(False)
            p = eng.profile_read()
            line += " | " + " ".join(f"{k}={v[0]/ms*100:.1f}%" for k, v in p.items() if v[2])
        print(line, flush=True)
