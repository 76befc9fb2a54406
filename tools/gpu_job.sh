for d in 0 1 7 15 31 63 127 255 ; do VMLMF_DBG=$d python tools/time_fwd.py 2>&1 | grep -E "loop|fwd-train" | sed "s/^/dbg=$d /"; done
VMLMF_R1_SIMT=1 python tools/time_fwd.py 2>&1 | grep -E "loop|fwd-train|^train"
