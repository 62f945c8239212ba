"""A/B timing of the attention launches of one north_star forward under several library builds:
python tools/attn_ab.py variants/libcobevt_a.so variants/libcobevt_b.so ...   (each in a fresh process; COBEVT_ATTN_POLY passes through)"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for lib in sys.argv[1:]:
    for poly in os.environ.get("AB_POLY", "2").split(","):
        env = dict(os.environ, COBEVT_LIB_PATH=os.path.abspath(lib), COBEVT_ATTN_POLY=poly)
        out = subprocess.run([sys.executable, os.path.join(here, "profile_launches.py")], env=env, capture_output=True, text=True).stdout
        att = [l for l in out.splitlines() if "window_attn" in l]
        tot = [l for l in out.splitlines() if l.startswith("total")]
        us = [float(l.split()[0]) for l in att]
        print(f"{os.path.basename(lib):28s} poly={poly} attn1 {us[0]:7.1f} attn2 {us[1]:6.1f} s1 {us[2]:6.1f} s2seg {us[4]:6.1f} glob {us[6]:6.1f} "
              f"fuse {us[7]:5.1f} | attn sum {sum(us):7.1f} us | {tot[0] if tot else ''}", flush=True)
