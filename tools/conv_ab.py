"""A/B timing of the GEMM-family launches of one north_star forward under several library builds:
python tools/conv_ab.py default variants/libcobevt_x.so ...   ('default' = the in-tree library)"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["COBEVT_LIB_PATH"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, os.path.join(here, "profile_launches.py")], env=env, capture_output=True, text=True).stdout
    rows = [l.split() for l in out.splitlines() if " conv_gemm " in l]
    us = [float(r[0]) for r in rows]
    enc = sum(us[:29])
    tot = [l for l in out.splitlines() if l.startswith("total")]
    print(f"{os.path.basename(lib):24s} encoder(29) {enc:7.1f} us | all gemm-family {sum(us):7.1f} us ({len(us)} launches) | layer1 {us[1]:.1f} {us[2]:.1f} layer2 {us[9]:.1f} {us[10]:.1f} "
          f"layer3 {us[18]:.1f} {us[19]:.1f} | fproj {us[29]:.1f} q {us[33]:.1f} proj {us[34]:.1f} | {tot[0] if tot else ''}", flush=True)
