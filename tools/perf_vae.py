import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
cfg = dict(bench.WORKLOADS["wan2.1-t2v-14b-720p-81f"])
print(json.dumps(bench.vae_decode_bench(cfg, torch.device("cuda", 0), with_reference=True)))
