import importlib, json, os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, bench
pkg = importlib.import_module("recursive-stwo_b200"); pkg.init(0)
shape = pkg.shape_from_config(pkg.PcsConfig(0, 5, 2, 16), 4, 8)
print(json.dumps(bench.synthetic_leg(pkg, torch.device("cuda:0"), shape)))
