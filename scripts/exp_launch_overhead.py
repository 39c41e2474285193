"""How long is one pomdp_rock_step launch when there is (almost) nothing to do?  Graph vs eager."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_pomdp_b200 as gp
dev = torch.device("cuda", 0)
for B in (1024, 1 << 16, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 24):
    env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=B, device=dev, seed=1)
    st, _ = env.init_states(B)
    a = torch.randint(0, 16, (B,), device=dev, dtype=torch.int32)
    out = (torch.empty_like(st), torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, device=dev), torch.empty(B, dtype=torch.int32, device=dev))
    K = 500
    for _ in range(5):
        env.simulate(st, a, out=out, step_ctr=1)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(K):
                env.simulate(st, a, out=out, step_ctr=i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay(); torch.cuda.synchronize()
        e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
        tg = e0.elapsed_time(e1) / K * 1e3
        e0.record(s)
        for i in range(K):
            env.simulate(st, a, out=out, step_ctr=i)
        e1.record(s); torch.cuda.synchronize()
        te = e0.elapsed_time(e1) / K * 1e3
    print("B=%9d  graph %.2f us/launch   eager %.2f us/launch   (%.1f MB/launch; L2-resident if < 126 MB)" % (B, tg, te, B * 24 / 1e6))
