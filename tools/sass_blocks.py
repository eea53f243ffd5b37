"""SASS block lengths of the default marcher, for bench.py's issue-slot roofline.

    python tools/sass_blocks.py [--lib volumeraytracer_b200/libvrt_b200.so] [--kernel MANGLED] [--sass-out FILE] > profiles/rNN_sass_blocks.json

The instrumented copy of the kernel (VRT_OPT_KERNEL 10) counts how often each block of march3_kernel is ISSUED (once per warp
pass).  This tool reads the lengths of those blocks -- in SASS instructions -- off the shipped binary: it disassembles the
production kernel (march3_kernel<float,false,false,false,9>) with nvdisasm, finds the loop nest from the backward branches and
classifies the blocks by what they contain:
    fast loop      innermost loop holding MUFU.RCP and the LDG.E.128 corner loads       -> `fast_step` (without the reload region)
    reload         the block inside the fast loop that holds the LDG.E.128 corner loads, from the branch that skips it to the
                   reconvergence BSYNC (exclusive)                                        -> `reload`; `reload_partial` = that BSYNC,
                   issued a second time when only a part of the active lanes took the reload
    for(;;) body   the loop around the fast loop (fast-loop exit checks + the straight-line generic step) -> `mid`, `generic`
    refill         the region of the outer loop that holds the ATOMG (ray counter)       -> `refill`
    retire         the region of the outer loop that holds the STG of the results        -> `retire`
    outer          what is left of the outer (poll) loop                                  -> `outer`
warp-instructions of a pass = sum(count[block] * length[block]); bench.py checks that model against ncu's smsp__inst_executed of
the committed capture (profiles/).  The JSON also carries a hash of the kernel's opcode stream so that a stale file is detected."""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_KERNEL = "_ZN3vrt13march3_kernelIfLb0ELb0ELb0ELi9EEEvNS_11MarchParamsE"
INS = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*$")
LABEL = re.compile(r"^(\.L_x_\d+):\s*$")


def disassemble(lib, kernel):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
        text = ""
        for c in cubins:
            text += subprocess.run(["nvdisasm", "-c", os.path.join(tmp, c)], capture_output=True, text=True, check=True).stdout
    start = text.find(".text.%s:" % kernel)
    if start < 0:
        raise SystemExit("kernel %s not found in %s" % (kernel, lib))
    end = text.find("//--------------------- .text.", start)
    return text[start:end if end > 0 else len(text)]


def parse(sass):
    """-> list of (addr, text), dict label -> addr of the next instruction"""
    ins, labels, pending = [], {}, []
    for line in sass.splitlines():
        m = LABEL.match(line)
        if m:
            pending.append(m.group(1))
            continue
        m = INS.match(line)
        if m:
            addr = int(m.group(1), 16)
            for l in pending:
                labels[l] = addr
            pending = []
            ins.append((addr, m.group(2).strip()))
    return ins, labels


def analyse(sass):
    ins, labels = parse(sass)
    # drop the trailing self-branch / padding after the last EXIT-reachable code: keep everything up to the first `BRA` to itself
    addr_index = {a: i for i, (a, _) in enumerate(ins)}

    def target(text):
        m = re.search(r"`\((\.L_x_\d+)\)", text)
        return labels.get(m.group(1)) if m else None

    loops = []          # (start_addr, end_addr) from backward branches
    for a, t in ins:
        if re.search(r"\bBRA\b", t):
            tg = target(t)
            if tg is not None and tg < a:
                loops.append((tg, a))
    loops = sorted(set(loops), key=lambda l: (l[1] - l[0]))

    def body(lo, hi):
        return [(a, t) for a, t in ins if lo <= a <= hi]

    def has(b, pat):
        return any(re.search(pat, t) for _, t in b)

    fast = next((l for l in loops if has(body(*l), r"MUFU\.RCP") and has(body(*l), r"LDG\.E\.128")), None)
    if fast is None:
        raise SystemExit("fast loop not found")
    forl = next((l for l in loops if l[0] <= fast[0] and l[1] >= fast[1] and l != fast), None)
    outer = next((l for l in reversed(loops) if has(body(*l), r"ATOMG") and has(body(*l), r"\bSTG\b")), None)
    if forl is None or outer is None:
        raise SystemExit("loop nest not recognised: %s" % loops)

    def ssy_regions(lo, hi):
        """BSSY..(label) regions inside [lo, hi]: (bssy_addr, addr of the BSYNC at the region's label)"""
        out = []
        for a, t in body(lo, hi):
            if t.startswith("BSSY") or " BSSY" in t:
                tg = target(t)
                if tg is not None and tg <= hi + 0x10:
                    # the BSYNC sits just before the label's first instruction
                    idx = addr_index.get(tg)
                    end = ins[idx - 1][0] if idx else tg
                    out.append((a, end))
        return out

    fb = body(*fast)
    # reload block: what the forward branch just before the corner loads skips -- from the instruction after that branch up to (not
    # including) the reconvergence BSYNC it jumps to.  The BSYNC itself is issued by every warp pass, and a SECOND time when only a
    # part of the active lanes took the reload (the two groups each issue it): counted separately by the instrumented kernel.
    ldg = [a for a, t in fb if re.search(r"LDG\.E\.128", t)]
    skips = [(a, target(t)) for a, t in fb if re.search(r"\bBRA\b", t) and target(t) and a < ldg[0] and max(ldg) < target(t) <= fast[1]]
    bra, tgt = max(skips, key=lambda x: x[0])
    reload_r = (bra + 0x10, tgt - 0x10)
    if not re.search(r"BSYNC", dict(ins)[tgt]):
        raise SystemExit("reload block does not end at a BSYNC: %s" % dict(ins)[tgt])
    n_fast_all = len(fb)
    n_reload = len(body(*reload_r))
    n_for = len(body(*forl))
    ob = body(*outer)
    # refill: the smallest BSSY-free span holding the ATOMG, delimited by the forward branch that skips it
    atom = next(a for a, t in ob if "ATOMG" in t)
    skip = [(a, target(t)) for a, t in ob if re.search(r"\bBRA\b", t) and target(t) and a < atom < target(t)]
    ref_lo, ref_hi = max(skip, key=lambda s: s[0])          # innermost forward branch over the atomic ...
    wide = min(skip, key=lambda s: s[0])                    # ... and the outermost one (the whole `if (!exhausted)` block)
    n_refill = len(body(wide[0] + 0x10, wide[1] - 0x10))
    # retire: from the first instruction after the for-loop's reconvergence up to the outer loop's back edge, containing the STGs
    stg = [a for a, t in ob if re.search(r"\bSTG\b", t)]
    skip_r = [(a, target(t)) for a, t in ob if re.search(r"\bBRA\b", t) and target(t) and a > forl[1] and a < stg[0] < target(t)]
    ret_lo = max(skip_r, key=lambda s: s[0])[0] if skip_r else stg[0]
    n_retire = len(body(ret_lo + 0x10, max(stg) + 0x10))
    n_outer_all = len(ob)
    n_outer = n_outer_all - n_for - n_refill - n_retire
    # for(;;) body outside the fast loop: the generic step is the part holding the second MUFU.RCP / the F2I conversions
    rest = [(a, t) for a, t in body(*forl) if not (fast[0] <= a <= fast[1])]
    n_rest = len(rest)
    first_gen = next((a for a, t in rest if a > fast[1] and re.search(r"F2I|MUFU\.RCP|FFMA2|FMUL2", t)), None)
    n_mid = len([1 for a, t in rest if a < fast[0]]) + len([1 for a, t in rest if a > fast[1] and (first_gen is None or a < first_gen)])
    n_mid = min(n_mid, n_rest)
    ops = "\n".join(re.sub(r"\s+", " ", t) for _, t in ins)
    counts = {
        "fast_step": n_fast_all - n_reload, "reload": n_reload, "reload_partial": 1, "mid": n_mid, "generic": n_rest - n_mid,
        "refill": n_refill, "retire": n_retire, "outer": max(n_outer, 0),
    }
    cost_fast_all = sum(issue_cost(t) for _, t in fb)
    cost_reload = sum(issue_cost(t) for _, t in body(*reload_r))
    costs = {"fast_step": round(cost_fast_all - cost_reload, 2), "reload": round(cost_reload, 2), "reload_partial": 1.0}
    for nm in ("mid", "generic", "refill", "retire", "outer"):
        costs[nm] = round(counts[nm] * 1.4, 2)                 # not on the hot path: the kernel-wide average cost per instruction
    mix = {}
    for _, t in fb:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
        mix[op] = mix.get(op, 0) + 1
    return {
        "kernel": None, "instructions_total": len(ins), "blocks": counts, "blocks_issue_cycles": costs,
        "loops": {"outer": [hex(outer[0]), hex(outer[1])], "for": [hex(forl[0]), hex(forl[1])], "fast": [hex(fast[0]), hex(fast[1])],
                  "reload": [hex(reload_r[0]), hex(reload_r[1])]},
        "fast_loop_opcode_mix": dict(sorted(mix.items(), key=lambda kv: -kv[1])),
        "sass_sha1": hashlib.sha1(ops.encode()).hexdigest(),
        "stat_slots": {"outer": 0, "refill": 1, "fast_step": 2, "reload": 3, "mid": 4, "generic": 5, "retire": 6, "lane_steps": 7, "reload_partial": 8},
    }


# Scheduler cycles per warp-instruction, measured with tools/pipe_bench.cu on a B200 at the marcher's occupancy (8 warps per scheduler;
# profiles/r02_pipe_bench.jsonl): scalar fp32 ~1.1, packed fp32x2 2.0, logic / compare / permute / shift / IMAD 2.0, IADD3 1.0; a mix of two
# kinds costs (nearly) the SUM of its parts (FFMA2+LOP3 4.0 per pair, FFMA+LOP3 2.76 instead of 3.15) -- the pipes hardly overlap.  Kinds
# that were not measured (branches, loads, MOV, conversions, MUFU) are counted as 1.
ISSUE_COST = [(r"^(FFMA2|FMUL2|FADD2)", 2.0), (r"^(FFMA|FADD|FMUL)\b", 1.1), (r"^(LOP3|ISETP|PRMT|SHF|SEL|PLOP3|VIMNMX|IMNMX|LEA|FSETP|FMNMX|POPC|FLO)", 2.0),
              (r"^IMAD", 2.0), (r"^(IADD3|VIADD|IADD)", 1.0)]


def issue_cost(text):
    op = re.sub(r"^@!?U?P\d+\s+", "", text)
    for pat, c in ISSUE_COST:
        if re.search(pat, op):
            return c
    return 1.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "volumeraytracer_b200", "libvrt_b200.so"))
    ap.add_argument("--kernel", default=DEFAULT_KERNEL)
    ap.add_argument("--sass-out", default=None, help="also write the kernel's SASS listing here")
    args = ap.parse_args()
    sass = disassemble(args.lib, args.kernel)
    if args.sass_out:
        with open(args.sass_out, "w") as f:
            f.write("// %s\n// nvdisasm -c of %s\n" % (args.kernel, os.path.relpath(args.lib, ROOT)))
            f.write(sass)
    res = analyse(sass)
    res["kernel"] = args.kernel
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
