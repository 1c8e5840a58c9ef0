"""Development probe (GPU): per-launch CUDA-event trace of one warmed MLICPP_L forward, aggregated by layer group."""
import sys, os, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/trace.tsv"
name = sys.argv[3] if len(sys.argv) > 3 else "MLICPP_L"
net = bench.seeded_model(name, "cuda:0").set_precision("bf16")
x = weights.synthetic_image(B, 1088, 1920, seed=2024, kind="rand").cuda()
for _ in range(3):
    net(x)
torch.cuda.synchronize()
net._trace = True
net(x)
torch.cuda.synchronize()
net.trace_dump(out_path)
net._trace = False
rows = [l.rstrip("\n").split("\t") for l in open(out_path)]
tot = sum(float(r[1]) for r in rows)
print(f"B={B}: {len(rows)} launches, {tot/1e3:.2f} ms in-stream")
def group(lab):
    k = lab.split(" ")[0]
    m = re.match(r"(g_a|g_s|h_a|h_s)\.", k)
    if m: 
        mm = re.match(r"(g_[as]\.\w+\.\d+)", k)
        return mm.group(1) if mm else m.group(1)
    m = re.match(r"([a-z_]+?)\.\d+", k)
    return m.group(1) if m else k
agg = collections.OrderedDict()
for lab, us in rows:
    g = group(lab); a = agg.setdefault(g, [0, 0.0]); a[0] += 1; a[1] += float(us)
for g, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/tot*100:6.2f}% {t/1e3:8.3f} ms x{c:4d} {g}")
print("--- top 40 launches")
for lab, us in sorted(rows, key=lambda r: -float(r[1]))[:40]:
    print(f"{float(us):9.1f} us  {lab}")
