"""Static look at the hot loops of rg_kernel in a cuobjdump -sass listing: DFMA count, operand-reuse hits
(an operand slot flagged .reuse by the previous DFMA and read again from the same register), other
instructions, and the stall counts ptxas encoded (control bits 105..108 of the 128-bit instruction word).
usage: cuobjdump -sass rg_b200/librg_b200.so > x.sass; python tools/sass_reuse.py x.sass [n_blocks]

NOTE (round 2): this tool counts a hit whenever the previous DFMA *of the block* flagged the slot, even with other
instructions in between.  The hardware only honours the flag between ADJACENT instructions (tools/dfma_lab), so the
figures printed here overstate the real hits (round 1's "90 of 130" was ~63); tools/sass_dump_loop.py and
tools/alloc_search.py count adjacent hits only."""
import re, sys
txt = open(sys.argv[1]).read().split('\n')
top = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ins = []   # (text, hi_word)
cur = None
for l in txt:
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/', l)
    if m:
        cur = [m.group(2).strip(), None]; ins.append(cur); continue
    m = re.match(r'\s*/\* (0x[0-9a-f]{16}) \*/', l)
    if m and cur is not None and cur[1] is None:
        cur[1] = int(m.group(1), 16)
blocks, b = [], []
for t, hi in ins:
    if re.search(r'\b(BRA|BSYNC|RET|CALL|EXIT|BRX|WARPSYNC|BSSY)\b', t):
        if b: blocks.append(b)
        b = []
    else:
        b.append((t, hi))
blocks.sort(key=lambda bl: -sum('DFMA' in t for t, _ in bl))
def ops(t):
    body = t.split(None, 1)[1] if ' ' in t else ''
    return [o.strip() for o in body.split(',')]
for bl in blocks[:top]:
    nd = hits = 0; stall = 0; stall_d = 0
    prev = None
    for t, hi in bl:
        s = ((hi >> 41) & 0xf) if hi is not None else 0   # bits 105..108 -> hi word bits 41..44
        stall += s
        o = ops(t)
        if 'DFMA' in t:
            nd += 1; stall_d += s
            if prev is not None:
                for k in range(1, min(len(o), len(prev))):
                    if prev[k].endswith('.reuse') and prev[k].replace('.reuse', '').lstrip('-') == o[k].replace('.reuse', '').lstrip('-'):
                        hits += 1; break
        if 'DFMA' in t:
            prev = o   # the reuse latch of a slot survives instructions that do not use it (ptxas flags across LDS)
    print("block: %d instr, %d DFMA, %d reuse hits, %d misses, stall-count sum %d (DFMA %d), LDS %d" %
          (len(bl), nd, hits, nd - hits, stall, stall_d, sum('LDS' in t for t, _ in bl)))
