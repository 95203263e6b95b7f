"""Prints host/GPU facts and CPU-RNG fingerprints (to check seeded init is identical on the GPU box)."""
import hashlib, os, platform, torch
def h(t): return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]
print("cpus", os.cpu_count(), platform.processor(), "threads", torch.get_num_threads())
try:
    print(open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0])
except Exception as e: print(e)
torch.manual_seed(1024)
print("linear", h(torch.nn.Linear(512, 512).weight), "emb", h(torch.nn.Embedding(512, 512).weight),
      "conv", h(torch.nn.Conv3d(64, 64, 3).weight), "randn", h(torch.randn(1000, 3)), "rand", h(torch.rand(4097)))
g = torch.Generator("cpu").manual_seed(7); print("gen", h(torch.randn([512, 32], generator=g)))
if torch.cuda.is_available():
    p = torch.cuda.get_device_properties(0)
    print(p.name, p.multi_processor_count, p.total_memory // 2**20, "MiB", "smem/blk optin", p.shared_memory_per_block_optin)
