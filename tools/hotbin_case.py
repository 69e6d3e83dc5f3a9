import sys, os; sys.path.insert(0, '.')
from opticalraytrace_b200 import abi, lib
lib.init(1)
st = lib.make_settings(nphotons=1<<32, use_bottle=False)
sc,_ = lib.build_scene(st, 'res', 843e-9)
for ub in (False, True):
    st.use_bottle = int(ub)
    job = lib.job_from_settings(st, 2)
    lib.trace(job, sc)
    img, lost, hist, tm = lib.trace(job, sc)
    print("use_bottle", ub, "%.3e rays/s" % ((1<<32)/tm.trace_seconds), "binned %.3f" % (hist[0,0]/(1<<32)), "hottest bin share %.4f" % (img.max()/img.sum()), "nonempty", int((img>0).sum()), "d2h %.3f ms" % (tm.d2h_seconds*1e3))
lib.finalize()
