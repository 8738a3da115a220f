#!/bin/bash
# Counts the Blackwell-specific SASS mnemonics of the shipped library (no GPU needed):
#   UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores, UTCBAR = tcgen05.commit
#   bash tools/sass_histogram.sh > profiles/sass_histogram_r02.txt
cd "$(dirname "$0")/.." || exit 1
LIB=vit-of-pytorch_b200/libvitb200.so
echo "cuobjdump -sass $LIB  ($(stat -c %s $LIB) bytes, sources at $(git rev-parse --short HEAD 2>/dev/null))"
cuobjdump -sass "$LIB" > /tmp/vitb_sass.txt
for m in UTCHMMA "UTCHMMA.2CTA" LDTM STTM UTMALDG UTMASTG UTCBAR "SYNCS.PHASECHK" "SYNCS.ARRIVE" FFMA2 FMUL2 FADD2 "MUFU.EX2"; do
  printf "%-16s %6d\n" "$m" "$(grep -c "$m" /tmp/vitb_sass.txt)"
done
echo
echo "per kernel (tcgen05.mma / tcgen05.ld / TMA load / TMA store):"
awk '/Function :/{name=$3} /UTCHMMA/{a[name]++} /LDTM/{b[name]++} /UTMALDG/{c[name]++} /UTMASTG/{d[name]++} END{for(k in a) printf "%5d %5d %5d %5d  %s\n", a[k], b[k], c[k], d[k], k}' /tmp/vitb_sass.txt \
  | sed -E 's/_ZN[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_]+_cu_[0-9a-f]+//; s/E14CUtensorMap.*//; s/EvPK.*//' | sort -k5 | cut -c1-110
