"""Timeline of one end-to-end Segmenter.__call__ (32 x 10 s, pinned inputs): the call is replayed step by step with CUDA
events on the sub-batch streams (same streams, buffers and order as Segmenter._run_jobs) so that every copy and kernel
phase gets a start / end time relative to the call's start; plus the device-resident time of the same sub-batches run
back to back and of the whole batch in one forward.   python tools/e2e_timeline.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM

sd = syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM)
seg = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", max_batch=32)
eng = seg._engine
g = torch.Generator().manual_seed(1)
wav = torch.randn(32, 160000, generator=g).pin_memory()
clips = [wav[i:i + 1] for i in range(32)]
for _ in range(4):
    seg(wav=clips)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    seg(wav=clips)
torch.cuda.synchronize()
print("e2e call: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))

bounds = seg._sub_batches(32)
streams = eng.side_streams(len(bounds))
cstreams = eng.copy_streams(len(bounds))
main = torch.cuda.current_stream()
thr_n, thr_m = np.float32(2.6), np.float32(0.8)


def one_call(record):
    ev = lambda: torch.cuda.Event(enable_timing=True)
    start = ev(); start.record(main)
    h0 = time.perf_counter()
    marks = []
    for k, (lo, hi) in enumerate(bounds):
        st, cst = streams[k], cstreams[k]
        st.wait_stream(main)
        with torch.cuda.stream(st):
            e = [ev() for _ in range(5)]
            e[0].record(st)
            rows = [r.reshape(-1) for r in clips[lo:hi]]
            wav_dev, n_dev = eng.upload(rows, [160000] * (hi - lo), 160000, k)
            e[1].record(st)
            hidden, _, _, _ = eng.forward(wav_dev, n_dev, thr_n, thr_m, segment=False, slot=k)
            e[2].record(st)
            cst.wait_event(e[2])
            with torch.cuda.stream(cst):
                hh, hp = eng.pool.array(tuple(hidden.shape))
                hp.copy_(hidden, non_blocking=True)
                e[4].record(cst)
            sg, cnt, feat = eng.segment_states(hidden, thr_n, thr_m, slot=k)
            cnt_pin = torch.empty(cnt.shape, dtype=torch.int32, pin_memory=True); cnt_pin.copy_(cnt, non_blocking=True)
            seg_pin = torch.empty(sg.shape, dtype=torch.int32, pin_memory=True); seg_pin.copy_(sg, non_blocking=True)
            e[3].record(st)
        marks.append((k, hi - lo, e, time.perf_counter() - h0))
    for k, n, e, _ in marks:
        e[3].synchronize(); e[4].synchronize()
    h1 = time.perf_counter() - h0
    if record:
        print("host: all sub-batches enqueued after %.3f ms; everything complete after %.3f ms" % (marks[-1][3] * 1e3, h1 * 1e3))
        for k, n, e, hq in marks:
            t = [start.elapsed_time(x) for x in e]
            print("  sub-batch %d (%2d rows, enqueued at %.3f ms host): upload %.3f-%.3f | encoder -%.3f | segmentation + table copy -%.3f | hidden D2H -%.3f"
                  % (k, n, hq * 1e3, t[0], t[1], t[2], t[3], t[4]))


for _ in range(3):
    one_call(False)
one_call(True)
one_call(True)

# device-resident: the three sub-batches back to back on one stream, and the whole batch at once
st = streams[0]
with torch.cuda.stream(st):
    devs = []
    for k, (lo, hi) in enumerate(bounds):
        devs.append((wav[lo:hi].to("cuda"), torch.full((hi - lo,), 160000, dtype=torch.int32, device="cuda")))
    whole = (wav.to("cuda"), torch.full((32,), 160000, dtype=torch.int32, device="cuda"))
    def run_subs():
        for k, (w, n) in enumerate(devs):
            eng.forward(w, n, thr_n, thr_m, slot=("t", k))
    def run_whole():
        eng.forward(whole[0], whole[1], thr_n, thr_m, slot=("t", "w"))
    for f, name in ((run_subs, "sub-batches %s back to back" % [hi - lo for lo, hi in bounds]), (run_whole, "one batch of 32")):
        for _ in range(4):
            f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(20):
            f()
        b.record(st); b.synchronize()
        print("device-resident, %s: %.3f ms" % (name, a.elapsed_time(b) / 20))
