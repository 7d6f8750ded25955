#!/bin/bash
# One line per kernel of libss4k.so: tcgen05 / TMEM / TMA instruction counts from the SASS (run on the build box, no GPU needed).
SO=${1:-sharkshark-4k_b200/csrc/libss4k.so}
cuobjdump -sass "$SO" 2>/dev/null | awk '
  /Function :/ { f=$3 }
  /UTCHMMA/ {a[f]++} /UTCQMMA|UTCIMMA/ {a2[f]++} /LDTM/ {b[f]++} /STTM/ {c[f]++} /UTMALDG/ {d[f]++} /UTMASTG/ {e[f]++} /UTCBAR/ {g[f]++}
  /UTMAPF|UTMACCTL/ {h[f]++} /R2UR/ {r[f]++} /SYNCS/ {s[f]++}
  END { for (k in r) printf "%-70s UTCHMMA %3d  LDTM %3d  STTM %3d  UTMALDG %3d  UTMASTG %3d  UTCBAR %3d  SYNCS %3d  R2UR %3d\n", k, a[k], b[k], c[k], d[k], e[k], g[k], s[k], r[k] }' | sort | c++filt
