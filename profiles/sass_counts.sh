#!/bin/bash
# Per-kernel SASS mnemonic counts of the built library (no GPU needed):
#     bash profiles/sass_counts.sh > profiles/r02_sass.txt
# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, IMAD.WIDE.U32 = the Philox multiplies,
# FFMA2/FADD2 = Blackwell packed FP32, LDG/STG.E.128 = 16 B global accesses, VOTE = ballot (bit packing).
SO=${1:-fiber_b200/_lib/libfiber_b200.so}
echo "# $(sha256sum $SO | cut -c1-16)  $SO  ($(git rev-parse --short HEAD 2>/dev/null))"
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; order[++n]=fn }
  fn != "" {
    if ($0 ~ /UBLKCP/) c[fn,"UBLKCP"]++
    if ($0 ~ /SYNCS/) c[fn,"SYNCS"]++
    if ($0 ~ /IMAD\.WIDE\.U32/) c[fn,"IMAD.WIDE.U32"]++
    if ($0 ~ /FFMA2/) c[fn,"FFMA2"]++
    if ($0 ~ /FADD2/) c[fn,"FADD2"]++
    if ($0 ~ /LDG\.E\.(128|ENL2\.128|.*\.128)/) c[fn,"LDG.128"]++
    if ($0 ~ /STG\.E\.(128|.*\.128)/) c[fn,"STG.128"]++
    if ($0 ~ /VOTE/) c[fn,"VOTE"]++
    if ($0 ~ /ATOMG|RED\./) c[fn,"ATOM/RED"]++
    if ($0 ~ /DMUL|DADD|DSETP|I2F\.F64/) c[fn,"FP64"]++
    if ($0 ~ /^ +\/\*[0-9a-f]+\*\//) c[fn,"instr"]++
  }
  END {
    split("instr UBLKCP SYNCS IMAD.WIDE.U32 FFMA2 FADD2 LDG.128 STG.128 VOTE ATOM/RED FP64", k, " ")
    for (i=1;i<=n;i++) { fn=order[i]; line=fn ":"; for (j=1;j<=11;j++) if (c[fn,k[j]]>0) line=line " " k[j] "=" c[fn,k[j]]; print line }
  }' | c++filt | sed 's/(fbr::WaveParams)//; s/(fbr::GatherParams[^)]*)//'
