import sys, json, torch
sys.path.insert(0, '/root/repo')
import bench
torch.cuda.set_device(0)
out = bench.extra_configs(torch.device('cuda', 0))
print(json.dumps(out.get('north_star_samples_per_step_sweep'), indent=1), out.get('sweep_error'))
