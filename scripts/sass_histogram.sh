#!/bin/bash
# Opcode evidence per object of the built library (no GPU needed): tensor-core (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld), TMA
# (UTMALDG = cp.async.bulk.tensor, .GATHER4 = tile::gather4, UBLKCP = cp.async.bulk), Ampere-style async copies (LDGSTS),
# mbarrier traffic (SYNCS), register re-partitioning (USETMAXREG), programmatic dependent launch (ACQBULK / PREEXIT).
cd "$(dirname "$0")/../dfmdock_b200/lib"
for o in *.o; do
  echo "== $o"
  cuobjdump -sass $o | grep -oE '\b(UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG(\.[0-9A-Z]+)*|UTMASTG|UBLKCP(\.[A-Z]+)*|LDGSTS(\.[A-Z0-9]+)*|SYNCS(\.[A-Z0-9]+)*|USETMAXREG(\.[A-Z_]+)*|ACQBULK|PREEXIT|REDUX(\.[A-Z]+)*|MUFU\.TANH(\.F16)?|HFMA2)\b' | sort | uniq -c | sort -rn | awk '{printf "   %7d %s\n", $1, $2}'
done
