import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
from radarslampy_b200 import _ffi, synthetic as S
res_m=0.0438; n=17; rb=int(87.5/res_m)
w=S.World(); raw,poses=S.make_sequence(n,res_m=res_m,world=w)
pi,feats,counts=S.sequence_pairs(n,w,poses,res_m,rb,k=200,max_features=256)
cfg=_ffi.default_config(); cfg.range_bins=rb; cfg.cart_res_m=2*res_m; cfg.dist_thr_px=0.5/(2*res_m)
cfg.max_pairs,cfg.max_frames,cfg.max_features,cfg.write_cart_f32=16,17,256,0
fe=_ffi.RadarFE(cfg); b=fe.new_batch()
res,nxt,corr=b.track(raw,pi,feats,counts)
print("nodes",res["clique_nodes"],"good",res["n_good"],"inl",res["n_inliers"])
st,_=b.klt_status()
for p in range(3):
    g=st[p,:counts[p]].astype(bool)
    a=feats[p,:counts[p]][g]; bb=nxt[p,:counts[p]][g]
    fe.reject_outliers(a,bb)
    t=time.perf_counter()
    for _ in range(5): m,ni,nodes=fe.reject_outliers(a,bb)
    print(p,"K'",len(a),"inl",ni,"nodes",nodes,"ms/call",(time.perf_counter()-t)/5*1e3)
