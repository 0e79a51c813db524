import os, sys, time, torch, numpy as np
import torch.distributed as dist
sys.path.insert(0, os.getcwd())
from sylber_b200 import Segmenter, segment_sharded
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
seg = Segmenter(model_ckpt=None, state_dict=syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM), device=f"cuda:{local}", max_batch=32)
wav = torch.randn(32, 160000, generator=torch.Generator().manual_seed(1 + rank)).pin_memory()
clips = [wav[i:i + 1] for i in range(32)]
for _ in range(4):
    segment_sharded(seg, wav=clips, local_input=True, pad_to=160000, per_rank=32)
dist.barrier(); torch.cuda.synchronize()
tt = []
t0 = time.perf_counter()
for _ in range(20):
    a = time.perf_counter()
    segment_sharded(seg, wav=clips, local_input=True, pad_to=160000, per_rank=32)
    from sylber_b200.distributed import LAST_TIMING as LT
    tt.append((time.perf_counter() - a, LT["setup"], LT["call"], LT["wait_gather"], LT["unpack"]))
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / 20
for _ in range(3):
    seg(wav=clips, pad_to=160000)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    seg(wav=clips, pad_to=160000)
torch.cuda.synchronize()
loc = (time.perf_counter() - t0) / 20
m = np.array(tt).mean(0) * 1e3
print(f"rank {rank}: sharded {tot*1e3:.3f} ms (setup {m[1]:.3f}, call_with_tables {m[2]:.3f}, wait for gather {m[3]:.3f}, unpack {m[4]:.3f}, outside {m[0]-m[1:].sum():.3f}) | local only {loc*1e3:.3f} ms", flush=True)
dist.destroy_process_group()
