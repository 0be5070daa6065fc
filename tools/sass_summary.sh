#!/bin/bash
# SASS evidence of the Blackwell-native paths (no GPU needed): mnemonic counts per object built by llmrec_b200/build.py
#   tcgen05.mma -> UTCHMMA ; tcgen05.ld/st -> LDTM/STTM ; TMA tensor loads -> UTMALDG ; cp.async.bulk -> UBLKCP ; legacy mma.sync -> " HMMA" (must be 0)
cd "$(dirname "$0")/../llmrec_b200/build"
for f in proj_tc2 proj_tc score_tc spmm; do
  cuobjdump -sass $f.o > /tmp/_$f.sass 2>/dev/null
  printf "%-10s" "$f.o"
  for m in UTCHMMA UTMALDG UBLKCP LDTM STTM "SYNCS.ARRIVE.TRANS64" "SYNCS.PHASECHK"; do printf " %s=%s" "$m" "$(grep -c "$m" /tmp/_$f.sass)"; done
  printf " legacy_HMMA=%s\n" "$(grep -E '[^C]HMMA' /tmp/_$f.sass | grep -vc UTCHMMA)"
  grep -E "Function : " /tmp/_$f.sass | sed 's/.*Function : /    /' | c++filt | cut -c1-110
done
