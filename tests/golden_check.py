"""Shared checker: replay one reference fixture through an engine and compare with the golden outputs."""
import ref_host


def _body(text):
    return [l for l in text.splitlines() if l and not l.startswith("#")]


def _info(line):
    out = {}
    for kv in line.split("\t")[7].split(";"):
        p = kv.split("=", 1)
        out[p[0]] = p[1] if len(p) == 2 else True
    return out


def _key(line):
    f = line.split("\t")
    return (f[0], f[1], f[3], f[4])


def _weight_str(w):
    # WithInfoEvalSynchronizer.java:137 String.format("%.3g", w) when |w - 1| > 0.001
    return None if abs(w - 1.0) <= 0.001 else format(w, "#.3g")


def bgzf_side_loader(library):
    """A side loader for run_vcfeval that goes the whole way of the device loaders: the VCF text is bgzipped and indexed here
    (oracle/formats_oracle.py), then the sequence's records are found through vce_tabix_query and read through
    vce_vcf_parse_bgzf (inflate + parse + --ref-overlap trimming on the engine under test)."""
    import formats_oracle as fo
    cache = {}

    def load(text, name, template_len, sample, ploidy, pass_only, max_length, ref_overlap, region=None):
        if id(text) not in cache:
            raw = text.encode()
            lines = text.split("\n")
            header_len = 0
            recs = []
            names = []
            pos = 0
            data, index = fo.bgzf_write(raw, 777)     # small members: records straddle them
            for ln in lines:
                nxt = pos + len(ln) + 1
                if ln and not ln.startswith("#"):
                    f = ln.split("\t")
                    if f[0] not in names:
                        names.append(f[0])
                    beg = int(f[1]) - 1
                    recs.append((names.index(f[0]), beg, beg + max(1, len(f[3])), fo.voffset(index, pos), fo.voffset(index, min(nxt, len(raw)))))
                pos = nxt
            # tabix needs the records of a sequence to be contiguous; the fixtures are sorted by sequence
            cache[id(text)] = (data, library.tabix_open(fo.tabix_build(recs, names)) if recs else None)
        data, idx = cache[id(text)]
        rng = None
        if region is not None:      # --region: another sequence reads nothing; this one is queried and filtered by position
            rng = (region[1], region[2]) if region[0] == name else "other sequence"
        q = None
        if idx is not None and rng != "other sequence":
            q = idx.query(name) if rng is None else idx.query(name, rng[0], rng[1])
        if q is None:      # nothing indexed for the sequence / region: what TabixLineReader would read is empty
            side, _ = library.vcf_parse(b"", name, template_len, sample, ploidy, pass_only, max_length, ref_overlap)
            return side
        side, _ = library.vcf_parse_bgzf(data, q[0], q[1], name, template_len, sample, ploidy, pass_only, max_length, True, ref_overlap,
                                         None, rng)
        return side
    return load


def check_scenario(engine, sc, cfg_over=None, side_loader=None):
    """Returns the list of output names that were compared."""
    rr = ref_host.run_vcfeval(engine, sc["template_fa"], sc["baseline_vcf"], sc["calls_vcf"], sc["args"], cfg_over, side_loader)
    check_scenario.device_loaded = getattr(check_scenario, "device_loaded", 0) + getattr(rr, "device_loaded", 0)
    exp = sc["expected"]
    checked = []
    p = rr.params
    if "weighted_roc.tsv" in exp:
        assert rr.roc_tsv == exp["weighted_roc.tsv"], "weighted_roc.tsv differs"
        checked.append("weighted_roc.tsv")
    if p.output_mode == "split":
        sets = {"tp.vcf": [], "fp.vcf": [], "tp-baseline.vcf": [], "fn.vcf": []}
        for rec, cls, w, sync, st in rr.call_class:
            if cls == "TP":
                sets["tp.vcf"].append(_key(rec.line))
            elif cls in ("FP", "FP_CA"):
                sets["fp.vcf"].append(_key(rec.line))
        for rec, cls, sync, st in rr.base_class:
            if cls == "TP":
                sets["tp-baseline.vcf"].append(_key(rec.line))
            elif cls in ("FN", "FN_CA"):
                sets["fn.vcf"].append(_key(rec.line))
        for name, got in sets.items():
            if name in exp:
                want = [_key(l) for l in _body(exp[name])]
                assert got == want, "%s differs: %s vs %s" % (name, got, want)
                checked.append(name)
    if p.output_mode == "annotate":
        for side, fname, classes in (("calls", "calls.vcf", rr.call_class), ("baseline", "baseline.vcf", rr.base_class)):
            if fname not in exp:
                continue
            tag = "CALL" if side == "calls" else "BASE"
            known = {}
            for item in classes:
                rec = item[0]
                known[id(rec)] = item
            text = sc["calls_vcf"] if side == "calls" else sc["baseline_vcf"]
            _, recs = ref_host.parse_vcf(text)
            order = {}
            for r in recs:
                order.setdefault(r.chrom, []).append(r)
            got = []
            # known records are found through the loaded variants (same VcfRecord objects are not shared between
            # the two parses), so match by (chrom, refined ordinal)
            per_chrom = {}
            for item in classes:
                per_chrom.setdefault(item[0].chrom, {})[item[0].line] = item
            seq_names = [n for n, _ in ref_host.parse_fasta(sc["template_fa"])]
            for chrom in seq_names:
                for r in ref_host.sort_refine(order.get(chrom, [])):
                    item = per_chrom.get(chrom, {}).get(r.line)
                    if item is None:
                        got.append((_key(r.line), "IGN", None, None))
                    elif side == "calls":
                        _, cls, w, sync, st = item
                        wt = _weight_str(w) if cls in ("TP", "FP_CA") else None
                        got.append((_key(r.line), cls, str(sync) if sync is not None else None, wt))
                    else:
                        _, cls, sync, st = item
                        got.append((_key(r.line), cls, str(sync) if sync is not None else None, None))
            want = []
            for l in _body(exp[fname]):
                info = _info(l)
                want.append((_key(l), info.get(tag), info.get("SYNC"), info.get("CALL_WEIGHT")))
            assert got == want, "%s differs:\n got %s\nwant %s" % (fname, got, want)
            checked.append(fname)
    if p.output_mode == "combine" and "output.vcf" in exp:
        got = ref_host.format_combine(rr)
        want = _body(exp["output.vcf"])
        assert got == want, "output.vcf (combine) differs:\n" + "\n".join(
            "got  %s\nwant %s" % (g, w) for g, w in zip(got + [""] * len(want), want + [""] * len(got)) if g != w)
        checked.append("output.vcf")
    if p.output_mode == "ga4gh" and "output.vcf" in exp:
        got = ref_host.format_ga4gh(rr)
        want = _body(exp["output.vcf"])
        assert got == want, "output.vcf (ga4gh) differs:\n" + "\n".join(
            "got  %s\nwant %s" % (g, w) for g, w in zip(got + [""] * len(want), want + [""] * len(got)) if g != w)
        checked.append("output.vcf")
    if "summary.txt" in exp and p.output_mode in ("split", "combine"):
        assert rr.summary == exp["summary.txt"], "summary differs:\n%s\n%s" % (rr.summary, exp["summary.txt"])
        checked.append("summary.txt")
    if "phasing.txt" in exp:
        assert rr.phasing_txt == exp["phasing.txt"]
        checked.append("phasing.txt")
    if "expected_stderr" in sc and "too complex" in sc["expected_stderr"]:
        want = [l for l in sc["expected_stderr"].splitlines() if l.startswith("Evaluation too complex")]
        assert rr.warnings == want, "warnings differ: %s vs %s" % (rr.warnings, want)
        checked.append("stderr")
    return checked
