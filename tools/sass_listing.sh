#!/bin/bash
# SASS evidence (no GPU needed): per kernel, the counts of the mnemonics that prove the Blackwell-native paths
# (B200_PROFILING.md "What proves a Blackwell-native kernel").  Output: profiles/<round>_sass_mnemonics.md
R=${1:-r02}
SO=ngsdist_b200/libngsdist_b200.so
OUT=profiles/${R}_sass_mnemonics.md
cuobjdump -sass $SO > /tmp/ngsd_sass.txt
{
echo "# SASS mnemonics of \`$SO\` (sm_100a), \`cuobjdump -sass\`, $(date -u +%F)"
echo
echo "| kernel | instructions | UTCIMMA (tcgen05.mma i8) | STTM / LDTM (tcgen05.st / ld) | UBLKCP (cp.async.bulk, TMA bulk) | DMMA | IMMA | SYNCS (mbarrier) | DFMA+DMUL+DADD | LDS | POPC | REDUX/VOTE |"
echo "|---|---|---|---|---|---|---|---|---|---|---|---|"
python3 - <<'P'
import re, collections
cur=None; counts=collections.OrderedDict()
for line in open('/tmp/ngsd_sass.txt'):
    m=re.match(r'\s*Function : (\S+)', line)
    if m:
        cur=m.group(1); counts[cur]=collections.Counter(); continue
    m=re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op=m.group(1); counts[cur]['_n']+=1
        base=op.split('.')[0]
        counts[cur][base]+=1
import subprocess
def demangle(n):
    try: return subprocess.run(['c++filt', n], stdout=subprocess.PIPE, text=True).stdout.strip()
    except Exception: return n
for k,c in counts.items():
    name=demangle(k)
    name=re.sub(r'\(anonymous namespace\)::','',name); name=re.sub(r'\(.*','',name); name=name.replace('void ','')
    fp64=c['DFMA']+c['DMUL']+c['DADD']
    print("| `%s` | %d | %d | %d / %d | %d | %d | %d | %d | %d | %d | %d | %d |" % (name, c['_n'], c['UTCIMMA'], c['STTM'], c['LDTM'], c['UBLKCP'], c['DMMA'], c['IMMA'], c['SYNCS'], fp64, c['LDS'], c['POPC'], c['REDUX']+c['VOTE']))
P
echo
echo "No \`UTMALDG\` (tensor-map TMA) appears: operands are pre-tiled in HBM so that every stage is ONE contiguous \`cp.async.bulk\` per operand (\`UBLKCP\`)."
echo "FP64 has no \`tcgen05\` kind (ptxas rejects \`kind::f64\`), so the FP64 contraction is \`mma.sync.m8n8k4.f64\` = \`DMMA\`; the int8 contraction is \`tcgen05.mma.kind::i8\` = \`UTCIMMA\`."
} > $OUT
tail -30 $OUT
