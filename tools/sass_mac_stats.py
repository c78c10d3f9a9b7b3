#!/usr/bin/env python
"""Static look at the FFMA2 stream of a kernel in a cuobjdump -sass dump: how many FFMA2 of the scalar-tap form,
how many share the 64-bit sample operand with the instruction before (and carry .reuse), how many depend on the
instruction directly before (a serialised chain), plus counts of the non-FMA instructions.
usage: cuobjdump -sass X.o > x.sass; python tools/sass_mac_stats.py x.sass <substring of the kernel name>"""
import re
import sys


def main():
    txt = open(sys.argv[1]).read()
    pat = sys.argv[2]
    for fn in re.split(r'\n\s+Function : ', txt)[1:]:
        name = fn.split('\n')[0]
        if pat not in name:
            continue
        ins = re.findall(r'/\*[0-9a-f]{4,5}\*/\s+((?:@!?U?P\d+\s+)?[A-Z0-9_.]+[^;]*);', fn)
        ff = [(i, re.match(r'FFMA2 R(\d+), R(\d+)(\.reuse)?\.F32, R(\d+)(\.reuse)?\.F32x2\.HI_LO, (?:R(\d+)|RZ)', s)) for i, s in enumerate(ins)]
        ff = [(i, m) for i, m in ff if m]
        same_s = sum(1 for k in range(1, len(ff)) if ff[k][0] == ff[k - 1][0] + 1 and ff[k][1].group(4) == ff[k - 1][1].group(4))
        reuse_s = sum(1 for _, m in ff if m.group(5))
        reuse_t = sum(1 for _, m in ff if m.group(3))
        serial = sum(1 for k in range(1, len(ff)) if ff[k][0] == ff[k - 1][0] + 1 and ff[k][1].group(6) == ff[k - 1][1].group(1))
        ops = {}
        for s in ins:
            op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        print(name[-70:])
        print('  tap-form FFMA2 %d: same sample as the instruction before %d, sample .reuse %d, tap .reuse %d, back-to-back dependent %d'
              % (len(ff), same_s, reuse_s, reuse_t, serial))
        print('  ' + ' '.join('%s:%d' % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:24]))


if __name__ == '__main__':
    main()
