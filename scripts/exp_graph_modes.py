"""EXPERIMENT: torch.cuda.graph replay vs a graph captured by hand (cuda-python) vs eager, same buffers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda.bindings import runtime as rt
import gym_pomdp_b200 as gp
from gym_pomdp_b200 import _lib

dev = torch.device("cuda", 0)
B, n, k, K = 1 << 22, 11, 11, 2000
env = gp.make("Rock-v0", board_size=n, num_rocks=k, batch_size=B, device=dev, seed=0x5EED)
gen = torch.Generator(device=dev); gen.manual_seed(1)
sets = []
for _ in range(6):
    x = torch.randint(0, n, (B,), generator=gen, device=dev); y = torch.randint(0, n, (B,), generator=gen, device=dev)
    st = torch.randint(-1, 2, (B, k), generator=gen, device=dev)
    a = torch.randint(0, 5 + k, (B,), generator=gen, device=dev, dtype=torch.int32)
    s = env.pack(x, y, st)
    out = (torch.empty_like(s), torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, device=dev), torch.empty(B, dtype=torch.int32, device=dev))
    sets.append((s, a, out))
del x, y, st
def launch(i):
    s, a, o = sets[i % 6]
    env.simulate(s, a, out=o, step_ctr=i + 1)
for i in range(10): launch(i)
torch.cuda.synchronize()

def ev_time(fn, stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); fn(); e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3

stream = torch.cuda.Stream(dev)
# (a) eager
with torch.cuda.stream(stream):
    t = ev_time(lambda: [launch(i) for i in range(K)], stream)
print("eager              %.2f us/launch" % t)
# (b) torch graph
with torch.cuda.stream(stream):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for i in range(K): launch(i)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
print("torch.cuda.graph   %.2f us/launch" % ev_time(g.replay, stream))
print("torch.cuda.graph   %.2f us/launch" % ev_time(g.replay, stream))
# (c) manual capture on the same torch stream through cuda-python
h = stream.cuda_stream
with torch.cuda.stream(stream):
    err, = rt.cudaStreamBeginCapture(h, rt.cudaStreamCaptureMode.cudaStreamCaptureModeThreadLocal)
    assert err == 0, err
    for i in range(K): launch(i)
    err, graph = rt.cudaStreamEndCapture(h)
    assert err == 0, err
err, gexec = rt.cudaGraphInstantiate(graph, 0)
assert err == 0, err
err, nodes, nn = rt.cudaGraphGetNodes(graph)
print("manual graph nodes:", nn)
def replay_manual():
    rt.cudaGraphLaunch(gexec, h)
replay_manual(); torch.cuda.synchronize()
print("manual capture     %.2f us/launch" % ev_time(replay_manual, stream))
print("manual capture     %.2f us/launch" % ev_time(replay_manual, stream))
print("torch.cuda.graph   %.2f us/launch" % ev_time(g.replay, stream))
