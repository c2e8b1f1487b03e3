import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from radarslampy_b200 import _ffi, synthetic as S
res_m = 0.0438; rb = int(87.5/res_m); F = 256
import os
sys.path.insert(0, 'tools')
from seq_bench import cached_sequence
world = S.World(seed=1234)
raw, poses = S.make_sequence(F, res_m=res_m, world=world)
pair_idx, feats, counts = S.sequence_pairs(F, world, poses, res_m, rb, k=200, max_features=256)
def run(ext_stream, waits, NB=5, steps=40):
    cfg = _ffi.default_config(); cfg.range_bins = rb; cfg.cart_res_m = 2*res_m; cfg.dist_thr_px = 0.5/(2*res_m)
    cfg.max_pairs, cfg.max_frames, cfg.max_features = F-1, F, 256; cfg.write_cart_f32 = 0
    st = torch.cuda.Stream() if ext_stream else None
    fe = _ffi.RadarFE(cfg, device=0, stream=st.cuda_stream if st else None)
    bs = [fe.new_batch() for _ in range(NB)]
    for b in bs: b.upload(raw, pair_idx, feats, counts, prev_pose=poses[:-1], sync=False)
    fe.sync()
    def go(n):
        for i in range(n):
            if waits and i >= NB: bs[(i-NB) % NB].wait()
            bs[i % NB].run_async(with_mds=False)
    go(10); fe.sync()
    fe.timer_start(); go(steps); ms = fe.timer_stop_ms()
    print("ext_stream", ext_stream, "waits", waits, "ms/step", round(ms/steps, 3), flush=True)
    for b in bs: b.close()
    fe.close()
run(False, False); run(False, True); run(True, False); run(True, True)
