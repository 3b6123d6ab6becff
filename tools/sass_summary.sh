#!/bin/bash
# Counts of the SASS mnemonics that prove which hardware paths librqp.so uses (tcgen05 MMA, TMEM loads, TMA tensor
# and bulk copies, DMMA, cluster / multicast forms), written to profiles/sass_summary.txt.
cd "$(dirname "$0")/.."
LIB=reluqp-py_b200/lib/librqp.so
OUT=${1:-profiles/sass_summary.txt}
cuobjdump -sass $LIB > /tmp/librqp.sass
{
  echo "# cuobjdump -sass $LIB | grep -c <mnemonic>   ($(date -u +%Y-%m-%dT%H:%MZ), $(git rev-parse --short HEAD 2>/dev/null))"
  for m in UTCHMMA UTCHMMA.2CTA UTCBAR LDTM STTM UTMALDG.2D UTMALDG UBLKCP.S.G UTMAPF DMMA.8 F2F.F64.F32 DFMA FFMA \
           SYNCS.ARRIVE.TRANS64 SYNCS.PHASECHK ACQBULK REDG ATOMG LDG.E.128 LDS.128 SHFL; do
    printf "%-22s %s\n" "$m" "$(grep -c -- "$m" /tmp/librqp.sass)"
  done
  echo "# kernels"
  grep -o "Function : [^ ]*" /tmp/librqp.sass | sed 's/Function : //' | c++filt | sed 's/(.*//' | sort | uniq -c | sort -rn
} > $OUT
cat $OUT | head -40
