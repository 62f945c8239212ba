#!/bin/bash
# Same-box A/B of whole-forward per-launch timings under several library builds / environment settings:
#   tools/ab_all.sh TAG "name|lib-or-default|ENV=V ENV2=V" ...   -> gpurun_out/ab_TAG_name.log (profile_launches output)
TAG=$1; shift
for spec in "$@"; do
  IFS='|' read -r name lib envs <<< "$spec"
  ( [ "$lib" != "default" ] && export COBEVT_LIB_PATH=$(realpath "$lib"); for e in $envs; do export "$e"; done
    timeout 200 python tools/profile_launches.py 4 > gpurun_out/ab_${TAG}_${name}.log 2>&1 )
  python - "$name" gpurun_out/ab_${TAG}_${name}.log <<'PY'
import sys
name, path = sys.argv[1], sys.argv[2]
fam = {}
for l in open(path):
    f = l.split()
    if len(f) > 6 and f[1] == "us":
        key = "token_mlp" if "token_mlp" in l else f[6]
        fam[key] = fam.get(key, 0.0) + float(f[0])
tot = [l.strip() for l in open(path) if l.startswith("total")]
print(f"{name:14s}", " ".join(f"{k} {v:8.1f}" for k, v in sorted(fam.items())), "|", tot[0] if tot else "NO TOTAL")
PY
done
