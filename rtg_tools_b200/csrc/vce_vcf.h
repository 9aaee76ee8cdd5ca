// vce_vcf.h -- SURVEY.md section 8(f) rank 3: VCF text -> the flat variant arrays of include/vce.h, on the device.
//
// Restates, per record of ONE sequence's stream (what VcfRecordTabixCallable reads for a sequence):
//   the sort refinement of records that share a start      vcf/VcfSortRefiner.java:91-108 (java.util.PriorityQueue order)
//   the loader filters, in order                            E/VcfRecordTabixCallable.java:91-141 (max length, beyond the
//                                                           template, factory says no, factory skips)
//   the sample factory                                      E/VariantFactory.java:127-196 (FILTER / FT, GT rules, ploidy)
//   allele strings and types                                vcf/VcfUtils.java:785-828, vcf/VariantType.java:107-190
//   the allele table                                        E/Allele.java:145-192 (only REF and the GT's alleles exist)
// `--ref-overlap` trimming (Variant.trimAlleles) and evaluation regions stay with the host, as SURVEY 8(a3) says.
//
// Written once as pass bodies (VCE_LAMBDA) over an Ops policy, like vce_final.h: vce_vcf.cu runs them as kernels + CUB
// scans, the test-only emulation as loops.  One thread per byte (line starts), per line (parse), per run of equal starts
// (refinement), per kept variant (fill).
#ifndef VCE_VCF_H_
#define VCE_VCF_H_

#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/vce.h"
#include "vce_device.h"

#if defined(__CUDACC__)
#define VCE_VCF_LAMBDA [=] __host__ __device__
#else
#define VCE_VCF_LAMBDA [=]
#endif

namespace vce {

enum : uint8_t { VK_KEPT = 0, VK_NOT_STREAM = 1, VK_IGNORED = 2, VK_SKIPPED = 3, VK_FATAL = 4 };
// why a record is VK_FATAL (the reference throws VcfFormatException and the run ends)
enum : uint8_t { VF_NONE = 0, VF_FIELDS = 1, VF_POS = 2, VF_SAMPLES = 3, VF_GT = 4, VF_GT_RANGE = 5, VF_ALT_IS_REF = 6, VF_GT_WIDE = 7 };

struct VcfConfig {
  int32_t sample;        // sample column (0-based)
  int32_t ploidy;        // --sample-ploidy
  int32_t passOnly;      // !--all-records
  int32_t maxLength;     // --Xmax-length, -1 = no limit
  int32_t templateLen;   // records that end beyond it are skipped, -1 = unknown
  int32_t nameLen;       // CHROM to accept (a tabix query returns one sequence's records); 0 = accept every record
  int32_t refOverlap;    // --ref-overlap: Variant.trimAlleles (Variant.java:93-190) on the loaded variants
  int32_t nEvalRegions;  // --evaluation-regions of this sequence: merged intervals, ascending; -1 = none
  int32_t scoreLen;      // --vcf-score-field: 0 = QUAL, else the length of scoreField, a FORMAT key of the sample (GQ by default in vcfeval)
  char scoreField[20];
  int32_t regionBeg, regionEnd;   // --region within the sequence, 0-based, end exclusive; regionEnd <= regionBeg = the whole sequence
  const int32_t* evalRegions;   // HOST pointer to [start, end) pairs (vcfDrive makes the device copy: VcfArrays::evalRegions)
  char name[232];
};

struct VcfLine {           // 44 bytes per line of text
  int32_t start, end;      // POS - 1, start + len(REF)
  int32_t nAlt;
  int32_t ntBytes;         // bases of the materialised, non-null alleles
  int8_t gt[4];
  uint8_t kind, fatal, phased, pad;
  int32_t refOff, altOff;  // offsets of REF / ALT within the line (the fill pass reads them again)
  int32_t altLen, refLen;
  int32_t nSlots;          // allele slots of the variant: nAlt + 2, or 2 + the non-symbolic ALTs in all-alts mode (gt[0] == -3)
};

struct VcfArrays {
  // inputs
  const uint8_t* text;
  int32_t len;
  VcfConfig cfg;            // host copy
  const VcfConfig* dCfg;    // the same where the passes run (read through a pointer: a by-value copy would sit on every thread's stack)
  // pass scratch (device / host), sized by vcfDrive as the counts become known
  int32_t* byteScan;       // [len / 32 + 2] line starts per 32-byte block, scanned
  int32_t* lineStart;      // [nLines + 1]
  VcfLine* lines;          // [nLines]
  int32_t* streamScan;     // [nLines + 1]
  int32_t* streamLine;     // [nStream]
  int32_t* order;          // [nStream] refined position -> stream index
  int32_t* heap;           // [nStream] scratch of the refinement
  int32_t* keptScan;       // [nStream + 1]
  int32_t* skipScan;       // [nStream + 1]
  int32_t* keptPos;        // [n] kept variant -> refined position
  int32_t* firstFatal;     // [1] smallest line index with a fatal record
  // outputs (vce_variants layout), sized after the counting passes
  int32_t* id;             // [n]
  uint8_t* phased;         // [n]
  uint8_t* statusIn;       // [n]
  int8_t* gt;              // [n * ploidy]
  int32_t* alleleOff;      // [n + 1] (scan scratch)
  int32_t* alleleOffOut;   // [n + 1] its copy inside the output block
  int32_t* ntBase;         // [n + 1]
  int32_t* alleleStart;    // [nSlots]
  int32_t* alleleEnd;      // [nSlots]
  uint8_t* alleleNull;     // [nSlots]
  int32_t* alleleNtOff;    // [nSlots + 1]
  uint8_t* nt;             // [nNt]
  uint32_t* rocMask;       // [n] RocFilter bits of the variant's record (vcfRocFields)
  const int32_t* evalRegions;   // device copy of the evaluation regions (pairs), nullptr = none
  int32_t nEvalRegions;
  // --ref-overlap (sized n + 1 when on): current locus of every kept variant, trim mode, cluster heads, prefix maxima
  int32_t* locS;
  int32_t* locE;
  uint8_t* trim;           // 0 = untrimmed, 1 = prefix first, 2 = suffix first (Variant.trimAlleles(boolean) :97-122)
  uint8_t* head;           // 1 = no earlier variant's locus reaches this one's start: a new cluster begins
  int32_t* blockMax;       // [n / 1024 + 2]
  double* qual;            // [n] QUAL as a double, NaN when missing
  // counts
  int32_t nLines, nStream, n, nSlots, nNt, nSkipped;
};

// ------------------------------------------------------------------------------------------------ small helpers
VCE_HD bool vcfIsSym(const uint8_t* s, int n) {   // VariantType.getSymbolicAlleleType :107-122 != null
  if (n <= 0) return false;
  const uint8_t f = s[0], l = s[n - 1];
  return f == '<' || l == '>' || f == '*' || l == '*' || f == '[' || f == ']' || l == '[' || l == ']';
}
VCE_HD int vcfBaseCode(uint8_t c) {   // DnaUtils.encodeString (upper-cased); -1 = illegal
  switch (c) {
    case 'N': case 'n': return 0;
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'T': case 't': return 4;
    default: return -1;
  }
}
VCE_HD bool vcfEq(const uint8_t* a, int na, const char* b, int nb) {
  if (na != nb) return false;
  for (int i = 0; i < na; ++i) if (a[i] != (uint8_t) b[i]) return false;
  return true;
}
// one element of a ';'-separated list that is neither PASS nor "." -> filtered (VcfRecord.java:273-297)
VCE_HD bool vcfListFails(const uint8_t* s, int n) {
  int a = 0;
  for (int i = 0; i <= n; ++i) {
    if (i == n || s[i] == ';') {
      const int m = i - a;
      const bool ok = (m == 4 && s[a] == 'P' && s[a + 1] == 'A' && s[a + 2] == 'S' && s[a + 3] == 'S') || (m == 1 && s[a] == '.');
      if (!ok) return true;
      a = i + 1;
    }
  }
  return false;
}
// k-th element of a separated list; returns false when there are fewer
VCE_HD bool vcfNth(const uint8_t* s, int n, uint8_t sep, int k, int* off, int* len) {
  int a = 0, idx = 0;
  for (int i = 0; i <= n; ++i) {
    if (i == n || s[i] == sep) {
      if (idx == k) { *off = a; *len = i - a; return true; }
      ++idx;
      a = i + 1;
    }
  }
  return false;
}
VCE_HD int vcfCount(const uint8_t* s, int n, uint8_t sep) {
  int c = 1;
  for (int i = 0; i < n; ++i) if (s[i] == sep) ++c;
  return c;
}

// ------------------------------------------------------------------------------------------------ one record
// Everything the loader decides about one line; the allele bytes themselves are written by vcfFillVariant.
VCE_HDN void vcfParseLine(const uint8_t* t, int a, int b, const VcfConfig& cfg, VcfLine& L) {
  L.start = L.end = 0; L.nAlt = 0; L.ntBytes = 0;
  L.gt[0] = L.gt[1] = L.gt[2] = L.gt[3] = -2;
  L.kind = VK_NOT_STREAM; L.fatal = VF_NONE; L.phased = 0; L.pad = 0;
  L.refOff = L.altOff = L.altLen = L.refLen = 0;
  L.nSlots = 0;
  if (b > a && t[b - 1] == '\r') --b;
  bool blank = true;
  for (int i = a; i < b; ++i) if (t[i] != ' ' && t[i] != '\t') { blank = false; break; }
  if (blank || t[a] == '#') return;
  // fields: 0 CHROM 1 POS 3 REF 4 ALT 6 FILTER 8 FORMAT 9+sample
  int fo[7], fl[7];
  for (int q = 0; q < 7; ++q) { fo[q] = 0; fl[q] = -1; }
  const int want[7] = {0, 1, 3, 4, 6, 8, 9 + cfg.sample};
  int field = 0, fs = a;
  for (int i = a; i <= b; ++i) {
    if (i == b || t[i] == '\t') {
      for (int q = 0; q < 7; ++q) if (want[q] == field) { fo[q] = fs; fl[q] = i - fs; }
      ++field;
      fs = i + 1;
    }
  }
  const int nFields = field;
  if (cfg.nameLen > 0 && !vcfEq(t + fo[0], fl[0] < 0 ? 0 : fl[0], cfg.name, cfg.nameLen)) return;   // another sequence's record
  L.kind = VK_IGNORED;
  if (nFields < 8) { L.kind = VK_FATAL; L.fatal = VF_FIELDS; return; }
  // POS
  {
    long v = 0;
    if (fl[1] <= 0 || fl[1] > 10) { L.kind = VK_FATAL; L.fatal = VF_POS; return; }
    for (int i = 0; i < fl[1]; ++i) {
      const int c = t[fo[1] + i];
      if (c < '0' || c > '9') { L.kind = VK_FATAL; L.fatal = VF_POS; return; }
      v = v * 10 + (c - '0');
    }
    if (v > 0x7fffffffL) { L.kind = VK_FATAL; L.fatal = VF_POS; return; }
    L.start = (int32_t) v - 1;
  }
  const uint8_t* ref = t + fo[2];
  const int refLen = fl[2];
  L.refOff = fo[2] - a; L.refLen = refLen;
  L.end = L.start + refLen;
  // --region: TabixLineReader hands on only the records that overlap it (an empty REF still covers its position); the others
  // are not part of the stream (they take no ordinal)
  if (cfg.regionEnd > cfg.regionBeg) {
    const int32_t e1 = L.end > L.start + 1 ? L.end : L.start + 1;
    if (!(L.start < cfg.regionEnd && e1 > cfg.regionBeg)) { L.kind = VK_NOT_STREAM; return; }
  }
  const uint8_t* alt = t + fo[3];
  int altLen = fl[3];
  int nAlt = 0;
  if (!(altLen == 1 && alt[0] == '.')) nAlt = vcfCount(alt, altLen, ',');
  else altLen = 0;
  L.altOff = fo[3] - a; L.altLen = altLen; L.nAlt = nAlt;
  // loader filter 1: longest allele string (VcfRecordTabixCallable.java:103-106)
  int longest = refLen;
  for (int k = 0; k < nAlt; ++k) {
    int o = 0, l = 0;
    vcfNth(alt, altLen, ',', k, &o, &l);
    longest = l > longest ? l : longest;
  }
  if (cfg.maxLength > -1 && longest > cfg.maxLength) { L.kind = VK_SKIPPED; return; }
  // loader filter 2: beyond the template (:107-112)
  if (cfg.templateLen >= 0 && L.end > cfg.templateLen) { L.kind = VK_SKIPPED; return; }
  if (cfg.sample < 0) {
    // ---- AllAlts.variant (VariantFactory.java:223-239): no sample column; every ALT that is a sequence becomes an allele
    // (pruneEmptyAlts :211-221 drops the symbolic ones from the table), no genotype
    if (cfg.passOnly) {
      const uint8_t* fil = t + fo[4];
      const int filLen = fl[4];
      if (!(filLen == 1 && fil[0] == '.') && vcfListFails(fil, filLen)) return;
    }
    if (nAlt == 0) return;
    bool pad = false;
    {
      const uint8_t c = refLen > 0 ? ref[0] : (uint8_t) '.';
      bool hasAlt = false, same = true;
      for (int k = 0; k < nAlt; ++k) {
        int o = 0, l = 0;
        vcfNth(alt, altLen, ',', k, &o, &l);
        if (l > 0 && alt[o] != '*') { hasAlt = true; if (alt[o] != c) same = false; }
        else if (l == 0) same = false;
      }
      pad = hasAlt && same;
    }
    L.pad = pad ? 1 : 0;
    const int from = pad ? 1 : 0;
    int ntBytes = 0, real = 0;
    for (int id = 0; id <= nAlt; ++id) {
      const uint8_t* sq = ref;
      int sn = refLen;
      if (id > 0) {
        int o = 0, l = 0;
        vcfNth(alt, altLen, ',', id - 1, &o, &l);
        sq = alt + o; sn = l;
      }
      if (vcfIsSym(sq, sn)) continue;
      for (int q = from; q < sn; ++q) if (vcfBaseCode(sq[q]) < 0) { L.kind = VK_SKIPPED; return; }
      ntBytes += sn > from ? sn - from : 0;
      if (id > 0) ++real;
    }
    if (real == 0) return;
    L.ntBytes = ntBytes;
    L.nSlots = 2 + real;
    L.gt[0] = -3;   // all-alts marker
    L.kind = VK_KEPT;
    return;
  }
  // ---- SampleVariants.variant (VariantFactory.java:127-147)
  const uint8_t* fmt = t + fo[5];
  const int fmtLen = fl[5] < 0 ? 0 : fl[5];
  const bool haveFmt = nFields > 8;
  const int nSamples = nFields > 9 ? nFields - 9 : 0;
  const uint8_t* smp = t + fo[6];
  const int smpLen = fl[6] < 0 ? 0 : fl[6];
  // key -> index within FORMAT
  int gtKey = -1, ftKey = -1;
  if (haveFmt) {
    const int nk = vcfCount(fmt, fmtLen, ':');
    for (int k = 0; k < nk; ++k) {
      int o = 0, l = 0;
      vcfNth(fmt, fmtLen, ':', k, &o, &l);
      if (l == 2 && fmt[o] == 'G' && fmt[o + 1] == 'T' && gtKey < 0) gtKey = k;
      if (l == 2 && fmt[o] == 'F' && fmt[o + 1] == 'T' && ftKey < 0) ftKey = k;
    }
  }
  if (cfg.passOnly) {
    const uint8_t* fil = t + fo[4];
    const int filLen = fl[4];
    if (!(filLen == 1 && fil[0] == '.') && vcfListFails(fil, filLen)) return;     // FILTER not PASS
    if (ftKey >= 0 && cfg.sample < nSamples) {                                       // sample FT
      int o = 0, l = 0;
      if (vcfNth(smp, smpLen, ':', ftKey, &o, &l) && vcfListFails(smp + o, l)) return;
    }
  }
  // getDefinedVariantGt :163-196
  if (cfg.sample >= nSamples) { L.kind = VK_FATAL; L.fatal = VF_SAMPLES; return; }
  if (gtKey < 0) return;
  int go = 0, gl = 1;
  const uint8_t dot = '.';   // a sample with fewer sub-fields than FORMAT has keys reads as missing
  const uint8_t* gts = &dot;
  if (vcfNth(smp, smpLen, ':', gtKey, &go, &gl)) gts = smp + go; else gl = 1;
  int gt[4];
  int ng = 0;
  bool phased = false;
  {
    int x0 = 0;
    for (int i = 0; i <= gl; ++i) {
      if (i == gl || gts[i] == '/' || gts[i] == '|') {
        if (i < gl && gts[i] == '|') phased = true;
        const int m = i - x0;
        int v = -1;
        if (!(m == 1 && gts[x0] == '.')) {
          if (m <= 0 || m > 6) { L.kind = VK_FATAL; L.fatal = VF_GT; return; }
          v = 0;
          for (int q = 0; q < m; ++q) {
            const int c = gts[x0 + q];
            if (c < '0' || c > '9') { L.kind = VK_FATAL; L.fatal = VF_GT; return; }
            v = v * 10 + (c - '0');
          }
        }
        if (ng < 4) gt[ng] = v;
        ++ng;
        x0 = i + 1;
      }
    }
  }
  for (int q = 0; q < ng && q < 4; ++q) if (gt[q] > nAlt) { L.kind = VK_FATAL; L.fatal = VF_GT_RANGE; return; }
  if (ng > 4) {   // more alleles than the ABI's gt[] holds: cannot be the expected ploidy (<= 4) either
    L.kind = VK_SKIPPED;
    return;
  }
  if (cfg.ploidy == 2 && ng == 1) { gt[1] = gt[0]; ng = 2; }
  if (ng != cfg.ploidy) { L.kind = VK_SKIPPED; return; }   // SkippedVariantException: counted
  // allele strings (VcfUtils.getAlleleStrings :806-828): padding base removed when every non-'*' ALT starts with REF's first base
  bool pad = false;
  {
    const uint8_t c = refLen > 0 ? ref[0] : (uint8_t) '.';
    bool hasAlt = false, same = true;
    for (int k = 0; k < nAlt; ++k) {
      int o = 0, l = 0;
      vcfNth(alt, altLen, ',', k, &o, &l);
      if (l > 0 && alt[o] != '*') {
        hasAlt = true;
        if (alt[o] != c) same = false;
      } else if (l == 0) {
        same = false;
      }
    }
    pad = hasAlt && same;
  }
  L.pad = pad ? 1 : 0;
  const uint8_t* r0 = ref + (pad ? 1 : 0);
  const int r0n = refLen - (pad ? 1 : 0);
  for (int k = 0; k < nAlt; ++k) {   // an ALT equal to REF is a format error, whatever the GT says
    int o = 0, l = 0;
    vcfNth(alt, altLen, ',', k, &o, &l);
    const bool strip = pad && l > 0 && alt[o] != '*';
    const uint8_t* s = alt + o + (strip ? 1 : 0);
    const int sn = l - (strip ? 1 : 0);
    bool eq = sn == r0n;
    for (int q = 0; eq && q < sn; ++q) if (s[q] != r0[q]) eq = false;
    if (eq) { L.kind = VK_FATAL; L.fatal = VF_ALT_IS_REF; return; }
  }
  bool defined = false;   // some GT allele > 0 is a non-SV variant (VariantType.getType(String,String) :169-190)
  for (int q = 0; q < ng && !defined; ++q) {
    if (gt[q] <= 0) continue;
    int o = 0, l = 0;
    vcfNth(alt, altLen, ',', gt[q] - 1, &o, &l);
    const bool strip = pad && l > 0 && alt[o] != '*';
    const uint8_t* s = alt + o + (strip ? 1 : 0);
    const int sn = l - (strip ? 1 : 0);
    const bool snp = r0n == 1 && sn == 1 && s[0] != '*';
    if (snp || !vcfIsSym(s, sn)) defined = true;
  }
  if (!defined) return;
  for (int q = 0; q < ng; ++q) if (gt[q] > 127) { L.kind = VK_FATAL; L.fatal = VF_GT_WIDE; return; }
  // Allele.getAlleles :145-192: REF and the GT's alleles; symbolic ones are null; an illegal base skips the record
  int ntBytes = 0;
  for (int id = 0; id <= nAlt; ++id) {
    bool inGt = id == 0;
    for (int q = 0; q < ng; ++q) if (gt[q] == id) inGt = true;
    if (!inGt) continue;
    const uint8_t* s;
    int sn;
    if (id == 0) { s = ref; sn = refLen; }
    else {
      int o = 0, l = 0;
      vcfNth(alt, altLen, ',', id - 1, &o, &l);
      s = alt + o; sn = l;
    }
    if (vcfIsSym(s, sn)) continue;
    const int from = pad ? 1 : 0;
    for (int q = from; q < sn; ++q) if (vcfBaseCode(s[q]) < 0) { L.kind = VK_SKIPPED; return; }
    ntBytes += sn > from ? sn - from : 0;
  }
  L.ntBytes = ntBytes;
  for (int q = 0; q < ng; ++q) L.gt[q] = (int8_t) gt[q];
  L.phased = phased ? 1 : 0;
  L.nSlots = nAlt + 2;
  L.kind = VK_KEPT;
  return;
}

// ---- ROC inputs of one kept variant: what RocContainer.addRocLine (RocContainer.java:225-248) derives from the record
// VariantType of one ALT against REF, both with the padding base removed (VariantType.getType(String,String) :169-190,
// isInsertionOrDeletion :193-219).  Ordinals: 0 NO_CALL 1 UNCHANGED 2 SNP 3 MNP 4 DELETION 5 INSERTION 6 INDEL 7 SV_BREAKEND
// 8 SV_SYMBOLIC 9 SV_MISSING.
VCE_HD int vcfAlleleType(const uint8_t* r, int rn, const uint8_t* s, int sn) {
  bool eq = rn == sn;
  for (int i = 0; eq && i < rn; ++i) if (r[i] != s[i]) eq = false;
  if (eq) return 1;
  if (rn == 1 && sn == 1 && s[0] != '*') return 2;
  if (sn > 0) {
    const uint8_t f = s[0], l = s[sn - 1];
    if (f == '<' || l == '>') return 8;
    if (f == '*' || l == '*') return 9;
    if (f == '[' || f == ']' || l == '[' || l == ']') return 7;
  }
  if (rn == sn) return 3;
  bool insDel = rn == 0 || sn == 0;
  if (!insDel) {
    const bool ins = rn < sn;
    const uint8_t* s1 = ins ? r : s;   // the shorter one
    const uint8_t* s2 = ins ? s : r;
    const int n1 = ins ? rn : sn, n2 = ins ? sn : rn;
    int left = 0;
    while (left < n1 && s1[left] == s2[left]) ++left;
    if (left == n1) insDel = true;
    else {
      int right = 0;
      while (right < n1 && s1[n1 - right - 1] == s2[n2 - right - 1]) ++right;
      insDel = left + right >= n1;
    }
  }
  if (insDel) return rn < sn ? 5 : 4;
  return 6;
}
VCE_HD int vcfTypePrecedence(int a, int b) {   // VariantType.getPrecedence :85-98
  const bool ia = a >= 4 && a <= 6, ib = b >= 4 && b <= 6;
  if (a != b && ia && ib) return 6;
  return a > b ? a : b;
}
// Double.parseDouble for the decimal forms a QUAL column holds: [sign] digits [. digits] [e [sign] digits] with at most 15
// significant digits and a decimal exponent within +-22 is exact in one multiplication or division of exact doubles (both
// correctly rounded, like the JDK's result).  Anything else (hex floats, "NaN", "Infinity", more digits) is left to the
// caller's parser: *ok = false.
VCE_HD double vcfParseDecimal(const uint8_t* s, int n, bool* ok) {
  *ok = false;
  int i = 0;
  bool neg = false;
  if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
  uint64_t m = 0;
  int digits = 0, sig = 0, frac = 0;
  bool dot = false;
  for (; i < n; ++i) {
    const uint8_t c = s[i];
    if (c == '.' && !dot) { dot = true; continue; }
    if (c < '0' || c > '9') break;
    ++digits;
    if (sig > 0 || c != '0') { if (++sig > 15) return 0.0; m = m * 10 + (uint64_t) (c - '0'); }
    if (dot) ++frac;
  }
  if (digits == 0) return 0.0;
  int e10 = 0;
  if (i < n && (s[i] == 'e' || s[i] == 'E')) {
    ++i;
    bool eneg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) { eneg = s[i] == '-'; ++i; }
    int ed = 0;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; ++i) { if (++ed > 4) return 0.0; e10 = e10 * 10 + (s[i] - '0'); }
    if (ed == 0) return 0.0;
    if (eneg) e10 = -e10;
  }
  if (i != n) return 0.0;
  e10 -= frac;
  double v = (double) m;
  if (m != 0) {
    if (e10 > 22 || e10 < -22) return 0.0;
    double p = 1.0;
    for (int k = e10 < 0 ? -e10 : e10; k > 0; --k) p *= 10.0;   // exact up to 10^22
    v = e10 < 0 ? v / p : v * p;
  }
  *ok = true;
  return neg ? -v : v;
}
// Filter bits in RocFilter's declaration order (RocFilter.java:48-137): ALL HOM HET SNP NON_SNP MNP INDEL XRX NON_XRX; bit 31:
// the QUAL text needs the caller's own parser (qual is NaN then).  `line` is the record's text, L what vcfParseLine kept.
constexpr uint32_t kRocScoreUnparsed = 0x80000000u;
VCE_HDN void vcfRocFields(const uint8_t* line, int lineLen, const VcfLine& L, const VcfConfig& cfg, uint32_t* maskOut, double* qualOut) {
  const int ploidy = cfg.ploidy;
  // QUAL = field 5, INFO = field 7, FORMAT = field 8, the sample = field 9 + cfg.sample
  int fo[4] = {0, 0, 0, 0}, fl[4] = {-1, -1, -1, -1};
  int field = 0, fs = 0;
  const int last = cfg.scoreLen > 0 ? 9 + cfg.sample : 7;
  for (int i = 0; i <= lineLen && field <= last; ++i) {
    if (i == lineLen || line[i] == '\t') {
      if (field == 5) { fo[0] = fs; fl[0] = i - fs; }
      if (field == 7) { fo[1] = fs; fl[1] = i - fs; }
      if (field == 8) { fo[2] = fs; fl[2] = i - fs; }
      if (field == 9 + cfg.sample) { fo[3] = fs; fl[3] = i - fs; }
      ++field;
      fs = i + 1;
    }
  }
  uint32_t mask = 1u;   // ALL
  // zygosity (VcfUtils.isHomozygousAlt :692-694, isHeterozygous :708-710) on the sample's GT
  const bool hom = ploidy == 1 || L.gt[0] == L.gt[1];
  bool homRef = true;
  for (int q = 0; q < ploidy && q < 4; ++q) if (L.gt[q] != 0) homRef = false;
  if (hom && !homRef) mask |= 1u << 1;
  if (ploidy == 2 && L.gt[0] != L.gt[1]) mask |= 1u << 2;
  // VariantType.getType(VcfRecord, int[]) :130-160
  int type = 0, altId = 0, q = 0;
  bool allMissing = true;
  for (; q < ploidy && q < 4; ++q) {
    const int a = L.gt[q];
    if (a != -1) {
      allMissing = false;
      if (a != 0) { altId = a; ++q; break; }
    }
  }
  if (allMissing) type = 0;
  else if (altId == 0) type = 1;
  else {
    const int pad = L.pad ? 1 : 0;
    const uint8_t* ref = line + L.refOff + pad;
    const int rn = L.refLen - pad;
    const uint8_t* alt = line + L.altOff;
    auto typeOf = [&](int id) {
      int o = 0, l = 0;
      vcfNth(alt, L.altLen, ',', id - 1, &o, &l);
      const bool strip = pad && l > 0 && alt[o] != '*';
      return vcfAlleleType(ref, rn, alt + o + (strip ? 1 : 0), l - (strip ? 1 : 0));
    };
    type = typeOf(altId);
    for (; q < ploidy && q < 4; ++q) {
      const int a = L.gt[q];
      if (a > 0 && a != altId) type = vcfTypePrecedence(type, typeOf(a));
    }
  }
  mask |= type == 2 ? 1u << 3 : 1u << 4;
  if (type == 3) mask |= 1u << 5;
  if (type >= 4 && type <= 6) mask |= 1u << 6;
  // VcfUtils.isComplexScored :529-531: INFO holds the key XRX
  bool xrx = false;
  if (fl[1] > 0) {
    const uint8_t* info = line + fo[1];
    int a = 0;
    for (int i = 0; i <= fl[1]; ++i) {
      if (i == fl[1] || info[i] == ';') {
        int k = a;
        while (k < i && info[k] != '=') ++k;
        if (k - a == 3 && info[a] == 'X' && info[a + 1] == 'R' && info[a + 2] == 'X') xrx = true;
        a = i + 1;
      }
    }
  }
  mask |= xrx ? 1u << 7 : 1u << 8;
  // RocScoreField.QUAL :50-80
#if defined(__CUDA_ARCH__)
  double qual = __longlong_as_double(0x7ff8000000000000LL);   // Double.NaN
#else
  double qual = std::numeric_limits<double>::quiet_NaN();
#endif
  if (cfg.scoreLen == 0) {
    if (fl[0] > 0 && !(fl[0] == 1 && line[fo[0]] == '.')) {
      bool ok = false;
      const double v = vcfParseDecimal(line + fo[0], fl[0], &ok);
      if (ok) qual = v; else mask |= kRocScoreUnparsed;
    }
  } else if (fl[2] > 0 && fl[3] >= 0) {
    // RocScoreField.FORMAT :120-140 over VcfUtils.getDoubleFormatFieldFromRecord :550-566: the sample's value of the key, NaN when
    // the key or the value is missing, values within 1e-8 of zero become 0
    const uint8_t* fmt = line + fo[2];
    const int nk = vcfCount(fmt, fl[2], ':');
    int key = -1;
    for (int k = 0; k < nk && key < 0; ++k) {
      int o = 0, l = 0;
      vcfNth(fmt, fl[2], ':', k, &o, &l);
      if (vcfEq(fmt + o, l, cfg.scoreField, cfg.scoreLen)) key = k;
    }
    int o = 0, l = 0;
    if (key >= 0 && vcfNth(line + fo[3], fl[3], ':', key, &o, &l) && !(l == 1 && line[fo[3] + o] == '.') && l > 0) {
      bool ok = false;
      const double v = vcfParseDecimal(line + fo[3] + o, l, &ok);
      if (ok) qual = (v >= -0.00000001 && v <= 0.00000001) ? 0.0 : v; else mask |= kRocScoreUnparsed;
    }
  }
  *maskOut = mask;
  *qualOut = qual;
}

// ---- --ref-overlap: Variant.trimAlleles (Variant.java:93-190)
// ALT `id` of a kept variant as the allele table holds it (padding base removed); false when the slot is null (not in the GT, or
// symbolic).  REF: id 0.
VCE_HD bool vcfAlleleText(const uint8_t* line, const VcfLine& L, int ploidy, int id, const uint8_t** s, int* sn) {
  const int pad = L.pad ? 1 : 0;
  if (id == 0) { *s = line + L.refOff + pad; *sn = L.refLen - pad; return true; }
  bool have = L.gt[0] == -3;   // all-alts mode: every ALT
  for (int q = 0; q < ploidy && q < 4; ++q) if (L.gt[q] == id) have = true;
  if (!have) return false;
  int o = 0, l = 0;
  vcfNth(line + L.altOff, L.altLen, ',', id - 1, &o, &l);
  const uint8_t* a = line + L.altOff + o;
  if (vcfIsSym(a, l)) return false;
  *s = a + pad;
  *sn = l - pad;
  return true;
}
VCE_HD int vcfCommonPrefix(int offset, const uint8_t* a, int an, const uint8_t* b, int bn) {   // ByteUtils.longestPrefix on base codes
  const int n = (an < bn ? an : bn) - offset;
  int i = 0;
  while (i < n && vcfBaseCode(a[i]) == vcfBaseCode(b[i])) ++i;
  return i;
}
VCE_HD int vcfCommonSuffix(int offset, const uint8_t* a, int an, const uint8_t* b, int bn) {   // ByteUtils.longestSuffix
  const int n = (an < bn ? an : bn) - offset;
  int i = 0;
  while (i < n && vcfBaseCode(a[an - 1 - i]) == vcfBaseCode(b[bn - 1 - i])) ++i;
  return i;
}
// Bases trimmed off one ALT (Variant.trimAlleles(boolean) :97-122): prefix first, then the suffix of what is left -- or the other way
VCE_HD void vcfTrimOf(const uint8_t* ref, int rn, const uint8_t* alt, int an, bool leftFirst, int* lead, int* trail) {
  if (leftFirst) { *lead = vcfCommonPrefix(0, ref, rn, alt, an); *trail = vcfCommonSuffix(*lead, ref, rn, alt, an); }
  else { *trail = vcfCommonSuffix(0, ref, rn, alt, an); *lead = vcfCommonPrefix(*trail, ref, rn, alt, an); }
}
// Locus of a variant after trimming in the given mode: the bounds of its non-null ALTs (REF is null afterwards); also the bases left.
VCE_HD void vcfTrimmedLocus(const uint8_t* line, const VcfLine& L, int ploidy, bool leftFirst, int32_t* ls, int32_t* le, int32_t* ntBytes) {
  const uint8_t* ref; int rn;
  vcfAlleleText(line, L, ploidy, 0, &ref, &rn);
  const int32_t s0 = L.start + (L.pad ? 1 : 0), e0 = L.end;
  int32_t lo = 0x7fffffff, hi = -0x7fffffff - 1, bytes = 0;
  for (int id = 1; id <= L.nAlt; ++id) {
    const uint8_t* alt; int an;
    if (!vcfAlleleText(line, L, ploidy, id, &alt, &an)) continue;
    int lead, trail;
    vcfTrimOf(ref, rn, alt, an, leftFirst, &lead, &trail);
    lo = s0 + lead < lo ? s0 + lead : lo;
    hi = e0 - trail > hi ? e0 - trail : hi;
    bytes += an - lead - trail;
  }
  *ls = lo; *le = hi; *ntBytes = bytes;
}
// Variant.trimAlleles(List) :124-190 over one cluster [v0, v1) of kept variants (no variant outside it overlaps one inside).
// a.locS / a.locE hold the current loci (untrimmed at the start), a.trim receives every variant's mode.
VCE_HD const VcfLine& vcfKeptLine(const VcfArrays& a, int v, const uint8_t** line) {
  const int l = a.streamLine[a.order[a.keptPos[v]]];
  *line = a.text + a.lineStart[l];
  return a.lines[l];
}
VCE_HD bool vcfLociOverlap(int32_t as, int32_t ae, int32_t bs, int32_t be) { return bs < ae && be > as; }   // Range.overlaps
VCE_HDN void vcfTrimCluster(const VcfArrays& a, int ploidy, int v0, int v1) {
  int first = v0, last = v0;
  for (int vi = v0; vi < v1; ++vi) {
    const uint8_t* line;
    const VcfLine& L = vcfKeptLine(a, vi, &line);
    while (last < v1 && (last <= vi || vcfLociOverlap(a.locS[vi], a.locE[vi], a.locS[last], a.locE[last]))) ++last;
    while (first < vi && !vcfLociOverlap(a.locS[vi], a.locE[vi], a.locS[first], a.locE[first])) ++first;
    bool leftFirst = true;
    if (!(first == vi && last == vi + 1)) {
      // the longest prefix / suffix any ALT shares with REF: only when both exist does the neighbourhood decide (:141-186)
      const uint8_t* ref; int rn;
      vcfAlleleText(line, L, ploidy, 0, &ref, &rn);
      int lead = 0, trail = 0;
      for (int id = 1; id <= L.nAlt; ++id) {
        const uint8_t* alt; int an;
        if (!vcfAlleleText(line, L, ploidy, id, &alt, &an)) continue;
        const int p = vcfCommonPrefix(0, ref, rn, alt, an), q = vcfCommonSuffix(0, ref, rn, alt, an);
        lead = p > lead ? p : lead;
        trail = q > trail ? q : trail;
      }
      if (lead != 0 && trail != 0) {
        int lc = 0, rc = 0;
        for (int i = first; i < last; ++i) {
          if (i == vi) continue;
          const int lo = a.locE[i] - a.locS[vi], ro = a.locE[vi] - a.locS[i];
          if (lo <= 0 && ro <= 0) continue;
          if (lo > 0 && lo <= lead) ++lc;
          if (ro > 0 && ro < trail) ++rc;
        }
        leftFirst = lc > rc;
      }
    }
    a.trim[vi] = leftFirst ? 1 : 2;
    int32_t ls, le, bytes;
    vcfTrimmedLocus(line, L, ploidy, leftFirst, &ls, &le, &bytes);
    a.locS[vi] = ls;
    a.locE[vi] = le;
  }
}

// The allele table of one kept variant (slot 0 = the missing allele, slot 1 = REF, slot k + 1 = ALT k).
VCE_HDN void vcfFillVariant(const uint8_t* line, const VcfLine& L, int ploidy, int32_t slot0, int32_t nt0, int32_t* aStart, int32_t* aEnd,
                            uint8_t* aNull, int32_t* aNtOff, uint8_t* nt, int trim = 0) {
  const uint8_t* ref = line + L.refOff;
  const uint8_t* alt = line + L.altOff;
  int at = nt0;
  if (L.gt[0] == -3) {   // all-alts mode: slot 0 missing, slot 1 REF, then the ALTs that are sequences, closed up (pruneEmptyAlts)
    const uint8_t* r0; int rn;
    vcfAlleleText(line, L, ploidy, 0, &r0, &rn);
    int slot = 0;
    aNtOff[slot0] = at; aStart[slot0] = 0; aEnd[slot0] = 0; aNull[slot0] = 1;
    ++slot;
    aNtOff[slot0 + slot] = at;
    if (trim) { aStart[slot0 + slot] = 0; aEnd[slot0 + slot] = 0; aNull[slot0 + slot] = 1; }
    else {
      aStart[slot0 + slot] = L.start + (L.pad ? 1 : 0); aEnd[slot0 + slot] = L.end; aNull[slot0 + slot] = 0;
      for (int q = 0; q < rn; ++q) nt[at++] = (uint8_t) vcfBaseCode(r0[q]);
    }
    ++slot;
    for (int id = 1; id <= L.nAlt; ++id) {
      const uint8_t* s; int sn;
      if (!vcfAlleleText(line, L, ploidy, id, &s, &sn)) continue;
      int lead = 0, trail = 0;
      if (trim) vcfTrimOf(r0, rn, s, sn, trim == 1, &lead, &trail);
      aNtOff[slot0 + slot] = at;
      aStart[slot0 + slot] = L.start + (L.pad ? 1 : 0) + lead;
      aEnd[slot0 + slot] = L.end - trail;
      aNull[slot0 + slot] = 0;
      for (int q = lead; q < sn - trail; ++q) nt[at++] = (uint8_t) vcfBaseCode(s[q]);
      ++slot;
    }
    return;
  }
  if (trim) {   // --ref-overlap: REF becomes null, every ALT loses what it shares with REF (Variant.trimAlleles(boolean) :97-122)
    const uint8_t* r0; int rn;
    vcfAlleleText(line, L, ploidy, 0, &r0, &rn);
    for (int slot = 0; slot < L.nAlt + 2; ++slot) {
      const uint8_t* s; int sn;
      aNtOff[slot0 + slot] = at;
      if (slot < 2 || !vcfAlleleText(line, L, ploidy, slot - 1, &s, &sn)) {
        aStart[slot0 + slot] = 0; aEnd[slot0 + slot] = 0; aNull[slot0 + slot] = 1;
        continue;
      }
      int lead, trail;
      vcfTrimOf(r0, rn, s, sn, trim == 1, &lead, &trail);
      aStart[slot0 + slot] = L.start + (L.pad ? 1 : 0) + lead;
      aEnd[slot0 + slot] = L.end - trail;
      aNull[slot0 + slot] = 0;
      for (int q = lead; q < sn - trail; ++q) nt[at++] = (uint8_t) vcfBaseCode(s[q]);
    }
    return;
  }
  for (int slot = 0; slot < L.nAlt + 2; ++slot) {
    const int id = slot - 1;
    bool have = id == 0;
    for (int q = 0; q < ploidy && q < 4; ++q) if (id >= 0 && L.gt[q] == id) have = true;
    const uint8_t* s = ref;
    int sn = L.refLen;
    if (id > 0) {
      int o = 0, l = 0;
      vcfNth(alt, L.altLen, ',', id - 1, &o, &l);
      s = alt + o; sn = l;
    }
    aNtOff[slot0 + slot] = at;
    if (id < 0 || !have || vcfIsSym(s, sn)) {
      aStart[slot0 + slot] = 0; aEnd[slot0 + slot] = 0; aNull[slot0 + slot] = 1;
      continue;
    }
    aStart[slot0 + slot] = L.start + (L.pad ? 1 : 0);
    aEnd[slot0 + slot] = L.end;
    aNull[slot0 + slot] = 0;
    for (int q = L.pad ? 1 : 0; q < sn; ++q) nt[at++] = (uint8_t) vcfBaseCode(s[q]);
  }
}

VCE_HD void vcfAtomicMin(int32_t* p, int32_t v) {
#if defined(__CUDA_ARCH__)
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}

// ---- passes 1: line starts, parse, stream, refinement, kept / skipped counts.  Fills A.nLines, nStream, n, nSkipped, and
// per kept variant the slot / base counts scanned into alleleOff / ntBase (both sized n + 1 by the caller AFTER
// vcfCountLines; see vce_vcf.cu / the emulation for the calling order).
constexpr int kVcfBlock = 32;   // bytes of text per thread in the line-start passes

template <class Ops>
int vcfCountLines(Ops& ops, VcfArrays& A) {
  const VcfArrays a = A;
  const int nb = (a.len + kVcfBlock - 1) / kVcfBlock;
  // a line starts at byte i when i == 0 or byte i - 1 is a newline: counted per block of 32 bytes, scanned over blocks
  ops.parFor(nb + 1, VCE_VCF_LAMBDA(int b) {
    int c = 0;
    if (b < nb) {
      const int lo = b * kVcfBlock, hi = lo + kVcfBlock < a.len ? lo + kVcfBlock : a.len;
      for (int i = lo; i < hi; ++i) c += (i == 0 || a.text[i - 1] == '\n') ? 1 : 0;
    }
    a.byteScan[b] = c;
  });
  ops.exclusiveSum(a.byteScan, a.byteScan, nb + 1);
  A.nLines = ops.fetchInt(a.byteScan + nb);
  return A.nLines;
}

template <class Ops>
void vcfParsePasses(Ops& ops, VcfArrays& A) {
  const VcfArrays a = A;
  const int nLines = A.nLines;
  ops.parFor((a.len + kVcfBlock - 1) / kVcfBlock, VCE_VCF_LAMBDA(int b) {
    int at = a.byteScan[b];
    const int lo = b * kVcfBlock, hi = lo + kVcfBlock < a.len ? lo + kVcfBlock : a.len;
    for (int i = lo; i < hi; ++i) if (i == 0 || a.text[i - 1] == '\n') a.lineStart[at++] = i;
  });
  ops.parFor(1, VCE_VCF_LAMBDA(int) { a.lineStart[nLines] = a.len; *a.firstFatal = 0x7fffffff; });
  ops.parFor(nLines + 1, VCE_VCF_LAMBDA(int l) {
    if (l == nLines) { a.streamScan[l] = 0; return; }
    int e = a.lineStart[l + 1];
    if (e > a.lineStart[l] && a.text[e - 1] == '\n') --e;
    VcfLine& L = a.lines[l];
    vcfParseLine(a.text, a.lineStart[l], e, *a.dCfg, L);
    a.streamScan[l] = L.kind != VK_NOT_STREAM ? 1 : 0;
    if (L.kind == VK_FATAL) vcfAtomicMin(a.firstFatal, l);
  });
  ops.exclusiveSum(a.streamScan, a.streamScan, nLines + 1);
  A.nStream = ops.fetchInt(a.streamScan + nLines);
  const int nStream = A.nStream;
  ops.parFor(nLines, VCE_VCF_LAMBDA(int l) {
    if (a.lines[l].kind != VK_NOT_STREAM) a.streamLine[a.streamScan[l]] = l;
  });
  // VcfSortRefiner: records that share a start leave a java.util.PriorityQueue ordered by end (ties are NOT stable: the
  // binary heap's sift rules are replayed as they are).  One thread per run of equal starts; the run's slice of `heap`
  // is its queue.
  ops.parFor(nStream, VCE_VCF_LAMBDA(int s) {
    const int st = a.lines[a.streamLine[s]].start;
    if (s > 0 && a.lines[a.streamLine[s - 1]].start == st) return;
    int e = s + 1;
    while (e < nStream && a.lines[a.streamLine[e]].start == st) ++e;
    if (e - s == 1) { a.order[s] = s; return; }
    int32_t* q = a.heap + s;
    int n = 0;
    for (int x = s; x < e; ++x) {   // offer (siftUp)
      const int kx = a.lines[a.streamLine[x]].end;
      int k = n++;
      while (k > 0) {
        const int parent = (k - 1) >> 1;
        const int pe = q[parent];
        if (kx >= a.lines[a.streamLine[pe]].end) break;
        q[k] = pe;
        k = parent;
      }
      q[k] = x;
    }
    for (int out = s; out < e; ++out) {   // poll (siftDown)
      a.order[out] = q[0];
      const int x = q[--n];
      if (n > 0) {
        const int kx = a.lines[a.streamLine[x]].end;
        int k = 0;
        const int half = n >> 1;
        while (k < half) {
          int child = 2 * k + 1;
          int c = q[child];
          const int right = child + 1;
          if (right < n && a.lines[a.streamLine[c]].end > a.lines[a.streamLine[q[right]]].end) { child = right; c = q[child]; }
          if (kx <= a.lines[a.streamLine[c]].end) break;
          q[k] = c;
          k = child;
        }
        q[k] = x;
      }
    }
  });
  ops.parFor(nStream + 1, VCE_VCF_LAMBDA(int p) {
    const int kind = p < nStream ? (int) a.lines[a.streamLine[a.order[p]]].kind : (int) VK_NOT_STREAM;
    a.keptScan[p] = kind == VK_KEPT ? 1 : 0;
    a.skipScan[p] = kind == VK_SKIPPED ? 1 : 0;
  });
  ops.exclusiveSum(a.keptScan, a.keptScan, nStream + 1);
  ops.exclusiveSum(a.skipScan, a.skipScan, nStream + 1);
  A.n = ops.fetchInt(a.keptScan + nStream);
  A.nSkipped = ops.fetchInt(a.skipScan + nStream);
}

// ---- passes 2 (after the caller sized keptPos / alleleOff / ntBase by A.n): slot and base offsets
template <class Ops>
void vcfOffsetPasses(Ops& ops, VcfArrays& A) {
  const VcfArrays a = A;
  const int nStream = A.nStream, n = A.n;
  ops.parFor(nStream, VCE_VCF_LAMBDA(int p) {
    if (a.keptScan[p + 1] != a.keptScan[p]) a.keptPos[a.keptScan[p]] = p;
  });
  const int ploidy = A.cfg.ploidy;
  if (A.cfg.refOverlap && n > 0) {
    // Variant.trimAlleles(List) :124-190 looks at a variant's overlapping neighbours, earlier ones already trimmed: clusters of
    // transitively overlapping variants are independent of each other and are walked by one thread each.  Cluster heads: a
    // variant that starts at or after the largest end seen so far (prefix maximum in blocks of 1024).
    const int nb = (n + 1023) / 1024;
    ops.parFor(n, VCE_VCF_LAMBDA(int v) {
      const uint8_t* line;
      const VcfLine& L = vcfKeptLine(a, v, &line);
      a.locS[v] = L.start + (L.pad ? 1 : 0);
      a.locE[v] = L.end;
      a.trim[v] = 0;
    });
    ops.parFor(nb, VCE_VCF_LAMBDA(int b) {
      int32_t m = -0x7fffffff - 1;
      const int hi = (b + 1) * 1024 < n ? (b + 1) * 1024 : n;
      for (int v = b * 1024; v < hi; ++v) m = a.locE[v] > m ? a.locE[v] : m;
      a.blockMax[b] = m;
    });
    ops.parFor(1, VCE_VCF_LAMBDA(int) {   // exclusive prefix maximum over the blocks
      int32_t m = -0x7fffffff - 1;
      for (int b = 0; b < nb; ++b) { const int32_t x = a.blockMax[b]; a.blockMax[b] = m; m = x > m ? x : m; }
    });
    ops.parFor(nb, VCE_VCF_LAMBDA(int b) {
      int32_t m = a.blockMax[b];
      const int hi = (b + 1) * 1024 < n ? (b + 1) * 1024 : n;
      for (int v = b * 1024; v < hi; ++v) {
        a.head[v] = (v == 0 || a.locS[v] >= m) ? 1 : 0;
        m = a.locE[v] > m ? a.locE[v] : m;
      }
    });
    ops.parFor(n, VCE_VCF_LAMBDA(int v) {
      if (!a.head[v]) return;
      int e = v + 1;
      while (e < n && !a.head[e]) ++e;
      vcfTrimCluster(a, ploidy, v, e);
    });
  }
  const bool trimmed = A.cfg.refOverlap != 0;
  ops.parFor(n + 1, VCE_VCF_LAMBDA(int v) {
    if (v == n) { a.alleleOff[v] = 0; a.ntBase[v] = 0; return; }
    const uint8_t* line;
    const VcfLine& L = vcfKeptLine(a, v, &line);
    a.alleleOff[v] = L.nSlots;
    int32_t bytes = L.ntBytes;
    if (trimmed) { int32_t ls, le; vcfTrimmedLocus(line, L, ploidy, a.trim[v] == 1, &ls, &le, &bytes); }
    a.ntBase[v] = bytes;
  });
  ops.exclusiveSum(a.alleleOff, a.alleleOff, n + 1);
  ops.exclusiveSum(a.ntBase, a.ntBase, n + 1);
  A.nSlots = ops.fetchInt(a.alleleOff + n);
  A.nNt = ops.fetchInt(a.ntBase + n);
}

// ---- passes 3 (after the caller sized the allele arrays by A.nSlots / A.nNt): the arrays of vce_variants
template <class Ops>
void vcfFillPasses(Ops& ops, VcfArrays& A) {
  const VcfArrays a = A;
  const int n = A.n, ploidy = A.cfg.ploidy;
  const int nSlots = A.nSlots, nNt = A.nNt;
  const bool trimmed = A.cfg.refOverlap != 0;
  ops.parFor(n, VCE_VCF_LAMBDA(int v) {
    const int p = a.keptPos[v];
    const int l = a.streamLine[a.order[p]];
    const VcfLine& L = a.lines[l];
    a.id[v] = p + 1;   // 1-based ordinal of the record in the refined stream, counted before any filter (:100)
    a.phased[v] = L.phased;
    // VcfRecordTabixCallable.java:125-127: outside the evaluation regions (MergedIntervals.overlapped :172-179 on the variant's
    // untrimmed locus: the interval with the greatest start below the locus' end must reach beyond its start)
    uint8_t status = 0;
    if (a.evalRegions) {
      const int32_t vs = L.start + (L.pad ? 1 : 0), ve = L.end;
      int lo = 0, hi = a.nEvalRegions;          // first interval whose start is >= ve
      while (lo < hi) { const int m = (lo + hi) / 2; if (a.evalRegions[2 * m] < ve) lo = m + 1; else hi = m; }
      const bool over = lo > 0 && vs < a.evalRegions[2 * (lo - 1) + 1];
      if (!over) status = VCE_STATUS_OUTSIDE_EVAL;
    }
    a.statusIn[v] = status;
    if (a.cfg.sample >= 0) for (int q = 0; q < ploidy; ++q) a.gt[(size_t) v * ploidy + q] = q < 4 ? L.gt[q] : (int8_t) -2;
    vcfFillVariant(a.text + a.lineStart[l], L, ploidy, a.alleleOff[v], a.ntBase[v], a.alleleStart, a.alleleEnd, a.alleleNull,
                   a.alleleNtOff, a.nt, trimmed ? a.trim[v] : 0);
    int e = a.lineStart[l + 1];
    if (e > a.lineStart[l] && a.text[e - 1] == '\n') --e;
    if (e > a.lineStart[l] && a.text[e - 1] == '\r') --e;
    if (a.cfg.sample >= 0) vcfRocFields(a.text + a.lineStart[l], e - a.lineStart[l], L, *a.dCfg, a.rocMask + v, a.qual + v);
    else { a.rocMask[v] = 0; a.qual[v] = 0.0; }   // all-alts mode has no sample to classify: the ROC inputs are the caller's
  });
  ops.parFor(n + 1, VCE_VCF_LAMBDA(int v) { a.alleleOffOut[v] = a.alleleOff[v]; });
  ops.parFor(1, VCE_VCF_LAMBDA(int) { a.alleleNtOff[nSlots] = nNt; });
}

// ------------------------------------------------------------------------------------------------ driver (host code)
struct VcfStats {
  int64_t lines = 0, records = 0, kept = 0, skipped = 0;
  double h2dMs = 0, kernelMs = 0, d2hMs = 0;
  int32_t launches = 0, fatalLine = -1, fatalReason = 0;
};
struct VcfResult {   // what vce_vcf_variants hands out: ONE host block (page-locked in the CUDA library) laid out like the device's
  int32_t ploidy = 0, n = 0, nSlots = 0, nNt = 0;
  uint8_t* slab = nullptr;
  size_t slabBytes = 0;
  void (*release)(uint8_t*, size_t) = nullptr;
  int32_t *id = nullptr, *alleleOff = nullptr, *alleleStart = nullptr, *alleleEnd = nullptr, *alleleNtOff = nullptr;
  uint8_t *phased = nullptr, *statusIn = nullptr, *alleleNull = nullptr, *nt = nullptr;
  int8_t* gt = nullptr;
  uint32_t* rocMask = nullptr;
  double* qual = nullptr;
  VcfStats stats;
  VcfResult() = default;
  VcfResult(const VcfResult&) = delete;
  VcfResult& operator=(const VcfResult&) = delete;
  ~VcfResult() { if (slab && release) release(slab, slabBytes); }
};

// Offsets of the output arrays within one block (every array 16-byte aligned); the same layout on the device and on the host.
struct VcfLayout {
  size_t id, phased, statusIn, gt, alleleOff, alleleStart, alleleEnd, alleleNull, alleleNtOff, nt, rocMask, qual, total;
  VcfLayout(size_t n, size_t ploidy, size_t nSlots, size_t nNt) {
    size_t at = 0;
    auto take = [&at](size_t bytes) { const size_t o = at; at += (bytes + 15) & ~(size_t) 15; return o; };
    id = take(n * 4); phased = take(n); statusIn = take(n); gt = take(n * ploidy + 4); alleleOff = take((n + 1) * 4);
    alleleStart = take(nSlots * 4); alleleEnd = take(nSlots * 4); alleleNull = take(nSlots); alleleNtOff = take((nSlots + 1) * 4);
    nt = take(nNt + 4);
    rocMask = take(n * 4); qual = take(n * 8);
    total = at;
  }
};

// Ops: alloc<T>(n) / hostAlloc(bytes, &release) / upload / download / parFor / exclusiveSum / fetchInt / mark(k) (timing marks: 0 before the upload,
// 1 after it, 2 after the last kernel, 3 after the downloads).
template <class Ops>
int vcfDrive(Ops& ops, const uint8_t* text, int64_t len, const VcfConfig& cfg, VcfResult& R, std::string& err,
             const uint8_t* residentText = nullptr) {   // residentText: the text is in device memory already (vce_bgzf.h)
  if (len < 0 || len >= 0x7ffffff0LL) { err = "VCF text of 2 GB or more: hand it over in pieces (one sequence per call)"; return 1; }
  if (cfg.ploidy < 1 || cfg.ploidy > 4 || cfg.sample < -1) { err = "sample / ploidy out of range"; return 1; }
  R.ploidy = cfg.sample < 0 ? 0 : cfg.ploidy;   // all-alts mode (sample -1, VariantFactory.AllAlts): no genotypes
  VcfArrays A;
  std::memset(&A, 0, sizeof(A));
  A.len = (int32_t) len;
  A.cfg = cfg;
  ops.mark(0);
  uint8_t* dText = residentText ? nullptr : ops.template alloc<uint8_t>((size_t) len + 16);
  A.byteScan = ops.template alloc<int32_t>((size_t) len / kVcfBlock + 4);
  A.firstFatal = ops.template alloc<int32_t>(4);
  if ((!dText && !residentText) || !A.byteScan || !A.firstFatal) { err = "out of device memory for the VCF text"; return 2; }
  if (!residentText) ops.upload(dText, text, (size_t) len);
  A.text = residentText ? residentText : dText;
  VcfConfig* dCfg = ops.template alloc<VcfConfig>(1);
  if (!dCfg) { err = "out of device memory for the VCF text"; return 2; }
  ops.upload(dCfg, &A.cfg, sizeof(VcfConfig));
  A.dCfg = dCfg;
  ops.mark(1);
  const int nLines = vcfCountLines(ops, A);
  A.lineStart = ops.template alloc<int32_t>((size_t) nLines + 1);
  A.lines = ops.template alloc<VcfLine>((size_t) nLines + 1);
  A.streamScan = ops.template alloc<int32_t>((size_t) nLines + 1);
  A.streamLine = ops.template alloc<int32_t>((size_t) nLines + 1);
  A.order = ops.template alloc<int32_t>((size_t) nLines + 1);
  A.heap = ops.template alloc<int32_t>((size_t) nLines + 1);
  A.keptScan = ops.template alloc<int32_t>((size_t) nLines + 1);
  A.skipScan = ops.template alloc<int32_t>((size_t) nLines + 1);
  if (!A.lineStart || !A.lines || !A.streamScan || !A.streamLine || !A.order || !A.heap || !A.keptScan || !A.skipScan) {
    err = "out of device memory for the VCF line tables";
    return 2;
  }
  vcfParsePasses(ops, A);
  R.stats.lines = nLines;
  R.stats.records = A.nStream;
  R.stats.skipped = A.nSkipped;
  const int fatal = ops.fetchInt(A.firstFatal);
  if (fatal != 0x7fffffff) {
    VcfLine L;
    ops.download(&L, A.lines + fatal, sizeof(L));
    R.stats.fatalLine = fatal;
    R.stats.fatalReason = L.fatal;
    static const char* const why[] = {"", "fewer than 8 fields", "POS is not a number", "record does not contain enough samples",
                                      "GT is not a list of allele ids", "GT contains an allele id out of range",
                                      "an ALT allele is the same as REF", "allele id above 127"};
    err = std::string("VCF format error in line ") + std::to_string(fatal + 1) + ": " + why[L.fatal < 8 ? L.fatal : 0];
    return 3;
  }
  const int n = A.n;
  A.keptPos = ops.template alloc<int32_t>((size_t) n + 1);
  int32_t* offScan = ops.template alloc<int32_t>((size_t) n + 1);
  A.ntBase = ops.template alloc<int32_t>((size_t) n + 1);
  if (!A.keptPos || !offScan || !A.ntBase) { err = "out of device memory for the variants"; return 2; }
  if (cfg.nEvalRegions >= 0 && cfg.evalRegions) {
    int32_t* dReg = ops.template alloc<int32_t>((size_t) cfg.nEvalRegions * 2 + 2);
    if (!dReg) { err = "out of device memory for the evaluation regions"; return 2; }
    ops.upload(dReg, cfg.evalRegions, (size_t) cfg.nEvalRegions * 2 * sizeof(int32_t));
    A.evalRegions = dReg;
    A.nEvalRegions = cfg.nEvalRegions;
  }
  if (cfg.refOverlap) {
    A.locS = ops.template alloc<int32_t>((size_t) n + 1);
    A.locE = ops.template alloc<int32_t>((size_t) n + 1);
    A.trim = ops.template alloc<uint8_t>((size_t) n + 1);
    A.head = ops.template alloc<uint8_t>((size_t) n + 1);
    A.blockMax = ops.template alloc<int32_t>((size_t) n / 1024 + 2);
    if (!A.locS || !A.locE || !A.trim || !A.head || !A.blockMax) { err = "out of device memory for the variants"; return 2; }
  }
  A.alleleOff = offScan;
  vcfOffsetPasses(ops, A);
  const VcfLayout lay((size_t) n, (size_t) cfg.ploidy, (size_t) A.nSlots, (size_t) A.nNt);
  uint8_t* dOut = ops.template alloc<uint8_t>(lay.total + 16);
  if (!dOut) { err = "out of device memory for the allele tables"; return 2; }
  A.id = reinterpret_cast<int32_t*>(dOut + lay.id);
  A.phased = dOut + lay.phased;
  A.statusIn = dOut + lay.statusIn;
  A.gt = reinterpret_cast<int8_t*>(dOut + lay.gt);
  A.alleleOffOut = reinterpret_cast<int32_t*>(dOut + lay.alleleOff);
  A.alleleStart = reinterpret_cast<int32_t*>(dOut + lay.alleleStart);
  A.alleleEnd = reinterpret_cast<int32_t*>(dOut + lay.alleleEnd);
  A.alleleNull = dOut + lay.alleleNull;
  A.alleleNtOff = reinterpret_cast<int32_t*>(dOut + lay.alleleNtOff);
  A.nt = dOut + lay.nt;
  A.rocMask = reinterpret_cast<uint32_t*>(dOut + lay.rocMask);
  A.qual = reinterpret_cast<double*>(dOut + lay.qual);
  vcfFillPasses(ops, A);
  ops.mark(2);
  R.slab = ops.hostAlloc(lay.total + 16, &R.release);
  R.slabBytes = lay.total + 16;
  if (!R.slab) { err = "out of host memory for the result"; return 2; }
  ops.download(R.slab, dOut, lay.total);
  ops.mark(3);
  R.n = n; R.nSlots = A.nSlots; R.nNt = A.nNt;
  R.id = reinterpret_cast<int32_t*>(R.slab + lay.id);
  R.phased = R.slab + lay.phased;
  R.statusIn = R.slab + lay.statusIn;
  R.gt = reinterpret_cast<int8_t*>(R.slab + lay.gt);
  R.alleleOff = reinterpret_cast<int32_t*>(R.slab + lay.alleleOff);
  R.alleleStart = reinterpret_cast<int32_t*>(R.slab + lay.alleleStart);
  R.alleleEnd = reinterpret_cast<int32_t*>(R.slab + lay.alleleEnd);
  R.alleleNull = R.slab + lay.alleleNull;
  R.alleleNtOff = reinterpret_cast<int32_t*>(R.slab + lay.alleleNtOff);
  R.nt = R.slab + lay.nt;
  R.rocMask = reinterpret_cast<uint32_t*>(R.slab + lay.rocMask);
  R.qual = reinterpret_cast<double*>(R.slab + lay.qual);
  R.stats.kept = n;
  return 0;
}

}  // namespace vce
#endif
