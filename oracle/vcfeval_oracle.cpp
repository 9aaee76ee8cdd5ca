// vcfeval_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A plain, sequential C++17 restatement of the reference's vcfeval matching
// path, class by class, so that the CUDA engine can be checked bit-for-bit.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product library
// (rtg_tools_b200/csrc) never links or calls it.
//
// Parity pinning: tests/test_oracle_golden.py replays the reference's own
// fixtures (test/com/rtg/vcf/eval/resources/*, bundled by
// tests/golden/make_golden.py) and unit vectors (PathTest, HaplotypePlaybackTest,
// OrientorTest, PhasingEvaluatorTest, RocContainerTest) through this oracle.
//
// Every function cites the reference file:line it follows; paths are relative
// to /root/reference, E/ = src/com/rtg/vcf/eval/.
//
// The reference cannot be compiled here (pure Java, no JDK in the image), so
// this is a "port" oracle, not oracle/_ref.

#include "../include/vce.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int MAXH = VCE_MAX_PLOIDY;

// ---------------------------------------------------------------- Allele ---
// E/Allele.java:47-119
struct Allele {
  int start = 0, end = 0;
  const uint8_t* nt = nullptr;
  int len = 0;
};

// E/Allele.java:95-119 (note the REVERSED base comparison at :111-117)
static int alleleCompare(const Allele* a, const Allele* b) {
  if (a == b) return 0;
  if (a->start != b->start) return a->start < b->start ? -1 : 1;
  if (a->end != b->end) return a->end < b->end ? -1 : 1;
  if (a->len != b->len) return a->len < b->len ? -1 : 1;
  for (int i = 0; i < a->len; ++i) {
    if (a->nt[i] != b->nt[i]) return b->nt[i] < a->nt[i] ? -1 : 1;
  }
  return 0;
}

// --------------------------------------------------------------- Variant ---
// E/Variant.java:49-91,213-243 ; E/GtIdVariant.java:43-80
struct OrientedVariant;
struct Variant {
  int id = 0;
  int start = 0, end = 0;   // Allele.getAlleleBounds, E/Allele.java:195-206
  bool phased = false;
  uint8_t status = 0;
  int inputIndex = 0;
  std::vector<const Allele*> alleles;  // index = allele id + 1, nullptr = Java null
  int gt[MAXH];
  int gtPloidy = 0;
  std::vector<OrientedVariant> orients[2];  // cached Orientor.orientations per pass
  bool orientsReady[2] = {false, false};

  int numAlleles() const { return (int) alleles.size() - 1; }         // Variant.java:263
  const Allele* allele(int alleleId) const {                           // Variant.java:241
    return alleleId < -1 ? nullptr : alleles[alleleId + 1];
  }
};

// E/OrientedVariant.java:42-75,171-196
struct OrientedVariant {
  Variant* v = nullptr;
  int alleleIds[MAXH];
  int nIds = 0;
  bool original = false;
  double weight = 0;
  int alleleId() const { return alleleIds[0]; }
  const Allele* allele(int hap) const { return v->allele(alleleIds[hap]); }
};

static OrientedVariant makeOv(Variant* v, bool original, std::initializer_list<int> ids) {
  OrientedVariant o;
  o.v = v;
  o.original = original;
  o.nIds = 0;
  for (int i : ids) o.alleleIds[o.nIds++] = i;
  return o;
}
static OrientedVariant makeOvArr(Variant* v, bool original, const int* ids, int n) {
  OrientedVariant o;
  o.v = v;
  o.original = original;
  o.nIds = n;
  for (int i = 0; i < n; ++i) o.alleleIds[i] = ids[i];
  return o;
}

// util/Permutation.java:43-107 -- all distinct permutations in lexicographic order
struct Permutation {
  std::vector<int> p;
  bool first = true;
  explicit Permutation(const int* seq, int n) : p(seq, seq + n) { std::sort(p.begin(), p.end()); }
  bool step() {
    if (first) { first = false; return true; }
    const int n = (int) p.size();
    if (n <= 1) return false;
    int j = n - 2;
    while (p[j] >= p[j + 1]) {
      if (--j < 0) return false;
    }
    int l = n - 1;
    while (p[j] >= p[l]) --l;
    std::swap(p[j], p[l]);
    int k = j + 1;
    l = n - 1;
    while (k < l) { std::swap(p[k], p[l]); ++k; --l; }
    return true;
  }
};

// ---------------------------------------------------------------- Orientor ---
// E/Orientor.java
static void altFill(Variant* v, std::vector<OrientedVariant>& o, int* alleles, int H, int i) {
  // Orientor.AltOrientor.fillAlleles :320-337
  if (i == H) {
    Permutation p(alleles, H);
    while (p.step()) o.push_back(makeOvArr(v, false, p.p.data(), H));
  } else {
    const bool explicitHalfCall = v->allele(-1) != nullptr;
    for (int j = -1; j <= alleles[i - 1]; ++j) {
      if (j == 0 && !explicitHalfCall) continue;
      alleles[i] = j;
      altFill(v, o, alleles, H, i + 1);
    }
  }
}

static std::vector<OrientedVariant> orientations(int kind, int H, Variant* v) {
  std::vector<OrientedVariant> out;
  switch (kind) {
    case VCE_ORIENT_PHASED:
    case VCE_ORIENT_PHASED_INVERT:
      // Orientor.PhasedOrientor.orientations :131-147
      if (v->phased) {
        if (kind == VCE_ORIENT_PHASED_INVERT) {
          int rev[MAXH];
          int i = v->gtPloidy;
          for (int k = 0; k < v->gtPloidy; ++k) rev[--i] = v->gt[k];
          out.push_back(makeOvArr(v, false, rev, v->gtPloidy));
        } else {
          out.push_back(makeOvArr(v, true, v->gt, v->gtPloidy));
        }
        return out;
      }
      // falls through to the unphased orientor (:145)
      [[fallthrough]];
    case VCE_ORIENT_UNPHASED:
      // Orientor.UnphasedOrientor.orientations :74-99
      if (H == 2) {
        if (v->gt[0] != v->gt[1]) {
          out.push_back(makeOv(v, true, {v->gt[0], v->gt[1]}));
          out.push_back(makeOv(v, false, {v->gt[1], v->gt[0]}));
        } else {
          out.push_back(makeOv(v, true, {v->gt[0], v->gt[0]}));
        }
      } else {
        Permutation p(v->gt, v->gtPloidy);
        while (p.step()) out.push_back(makeOvArr(v, false, p.p.data(), v->gtPloidy));
      }
      return out;
    case VCE_ORIENT_SQUASH_GT: {
      // Orientor.SQUASH_GT.orientations :166-187
      int alt[MAXH];
      for (int i = 0; i < v->gtPloidy; ++i) alt[i] = v->gt[i];
      std::sort(alt, alt + v->gtPloidy);
      int prev = 0;
      for (int i = 0; i < v->gtPloidy; ++i) {
        const int id = alt[i];
        if (id > prev && v->allele(id) != nullptr) {
          out.push_back(makeOv(v, false, {id}));
          prev = id;
        }
      }
      return out;
    }
    case VCE_ORIENT_ALLELE_GT: {
      // Orientor.ALLELE_GT.orientations :206-233
      const int a = v->gt[0], b = v->gt[1];
      const bool aVar = a > 0 && v->allele(a) != nullptr;
      const bool bVar = b > 0 && v->allele(b) != nullptr;
      const int la = aVar ? a : b;
      const int lb = bVar ? b : a;
      if (la == lb) {
        out.push_back(makeOv(v, false, {0, la}));
        out.push_back(makeOv(v, false, {la, 0}));
        out.push_back(makeOv(v, false, {la, la}));
      } else {
        out.push_back(makeOv(v, false, {0, la}));
        out.push_back(makeOv(v, false, {0, lb}));
        out.push_back(makeOv(v, false, {la, 0}));
        out.push_back(makeOv(v, false, {lb, 0}));
        out.push_back(makeOv(v, false, {la, lb}));
        out.push_back(makeOv(v, false, {lb, la}));
      }
      return out;
    }
    case VCE_ORIENT_HAPLOID_POP:
      // Orientor.HAPLOID_POP.orientations :250-256
      for (int i = 0; i < v->numAlleles() - 1; ++i) out.push_back(makeOv(v, false, {i + 1}));
      return out;
    case VCE_ORIENT_DIPLOID_POP: {
      // Orientor.DIPLOID_POP.orientations :274-289
      const bool explicitHalfCall = v->allele(-1) != nullptr;
      for (int i = 1; i < v->numAlleles(); ++i) {
        for (int j = -1; j < i; ++j) {
          if (j == 0 && !explicitHalfCall) continue;
          out.push_back(makeOv(v, false, {i, j}));
          out.push_back(makeOv(v, false, {j, i}));
        }
        out.push_back(makeOv(v, false, {i, i}));
      }
      return out;
    }
    case VCE_ORIENT_ALT: {
      // Orientor.AltOrientor.orientations :311-319
      int r[MAXH];
      for (int i = 1; i < v->numAlleles(); ++i) {
        r[0] = i;
        altFill(v, out, r, H, 1);
      }
      return out;
    }
  }
  return out;
}

// ------------------------------------------------- persistent linked list ---
// util/BasicLinkedListNode.java:40-58 (shared tails, O(1) size).  Java relies on
// the GC; here nodes are intrusively reference counted.
template <typename T>
struct Node {
  T value;
  Node* next;
  int size;
  int refs;
};
template <typename T>
struct NodePool {
  std::vector<Node<T>*> freeList;
  std::vector<std::unique_ptr<Node<T>[]>> blocks;
  Node<T>* alloc() {
    if (freeList.empty()) {
      const int B = 4096;
      blocks.emplace_back(new Node<T>[B]);
      Node<T>* blk = blocks.back().get();
      for (int i = 0; i < B; ++i) freeList.push_back(blk + i);
    }
    Node<T>* n = freeList.back();
    freeList.pop_back();
    return n;
  }
  Node<T>* push(const T& v, Node<T>* tail) {  // takes its own reference on tail
    Node<T>* n = alloc();
    n->value = v;
    n->next = tail;
    n->size = tail ? tail->size + 1 : 1;
    n->refs = 1;
    if (tail) ++tail->refs;
    return n;
  }
  static Node<T>* retain(Node<T>* n) { if (n) ++n->refs; return n; }
  void release(Node<T>* n) {
    while (n && --n->refs == 0) {
      Node<T>* nx = n->next;
      freeList.push_back(n);
      n = nx;
    }
  }
};

template <typename T>
static std::vector<T> toReversedList(Node<T>* head) {  // BasicLinkedListNode.toReversedList :66-76
  std::vector<T> v;
  if (!head) return v;
  v.reserve(head->size);
  for (Node<T>* n = head; n; n = n->next) v.push_back(n->value);
  std::reverse(v.begin(), v.end());
  return v;
}

// ------------------------------------------------------ HaplotypePlayback ---
// E/HaplotypePlayback.java
struct OracleError { int code; };

struct HaplotypePlayback {
  static constexpr int INVALID = -1;
  const uint8_t* tmpl = nullptr;
  int tmplLen = 0;
  int templatePosition = -1;     // :57
  int positionInAllele = INVALID;// :60
  int lastAlleleEnd = -1;        // :62
  const Allele* nextAllele = nullptr;  // :65
  bool finished = false;
  std::vector<const Allele*> alleles;  // FIFO :52 (front at index qHead)
  int qHead = 0;

  bool queueEmpty() const { return qHead >= (int) alleles.size(); }

  void add(const Allele* a) {  // :91-106
    if (a == nullptr) return;
    if (a->start == a->end && a->len == 0) return;
    lastAlleleEnd = a->end;
    if (nextAllele == nullptr) {
      nextAllele = a;
    } else {
      if (queueEmpty()) { alleles.clear(); qHead = 0; }
      alleles.push_back(a);
    }
  }
  bool isNew(const Allele* a) const { return a == nullptr || a->start >= lastAlleleEnd; }  // :112
  uint8_t nt() const {  // :120-125
    if (positionInAllele == INVALID) return templatePosition < tmplLen ? tmpl[templatePosition] : 0;
    return nextAllele->nt[positionInAllele];
  }
  bool hasNext() const { return templatePosition < tmplLen - 1; }  // :131
  void step() { if (hasNext()) next(); else finished = true; }     // :135-141
  bool isOnTemplate() const { return positionInAllele == INVALID; }  // :220
  bool wantsFutureVariantBases() const {  // :151-164
    if (nextAllele == nullptr) return true;
    if (positionInAllele != INVALID && positionInAllele < nextAllele->len - 1) return false;
    for (int i = qHead; i < (int) alleles.size(); ++i) {
      if (alleles[i]->len > 0) return false;
    }
    return true;
  }
  void next() {  // :170-214
    if (isOnTemplate()) {
      ++templatePosition;
      if (nextAllele != nullptr && nextAllele->start == templatePosition) positionInAllele = 0;
    } else {
      ++positionInAllele;
    }
    if (nextAllele != nullptr) {
      while (true) {
        if (positionInAllele != nextAllele->len) break;
        templatePosition = nextAllele->end;
        positionInAllele = INVALID;
        if (!queueEmpty()) {
          nextAllele = alleles[qHead++];
        } else {
          nextAllele = nullptr;
          break;
        }
        if (templatePosition < nextAllele->start) break;
        positionInAllele = 0;
        if (templatePosition != nextAllele->start) throw OracleError{VCE_ERR_UNSUPPORTED};  // "Out of order alleles" :208
      }
    }
  }
  void moveForward(int position) {  // :258-267
    if (!isOnTemplate()) throw OracleError{VCE_ERR_MOVE_IN_VARIANT};
    templatePosition = position - 1;
    next();
  }
  bool matches(const HaplotypePlayback& o) const { return finished || o.finished || nt() == o.nt(); }  // :316
  int compareTo(const HaplotypePlayback& that) const {  // :326-365
    const int position = templatePosition - that.templatePosition;
    if (position != 0) return position;
    if (nextAllele == nullptr) return that.nextAllele == nullptr ? 0 : -1;
    if (that.nextAllele == nullptr) return 1;
    const int current = alleleCompare(nextAllele, that.nextAllele);
    if (current != 0) return current;
    const int varPos = positionInAllele - that.positionInAllele;
    if (varPos != 0) return varPos;
    int i = qHead, j = that.qHead;
    const int ni = (int) alleles.size(), nj = (int) that.alleles.size();
    while (i < ni) {
      if (j >= nj) return 1;
      const int future = alleleCompare(alleles[i++], that.alleles[j++]);
      if (future != 0) return future;
    }
    if (j < nj) return -1;
    return 0;
  }
};

// ---------------------------------------------------------------- HalfPath ---
// E/HalfPath.java
struct InclEntry { OrientedVariant* ov; };
struct Pools {
  NodePool<OrientedVariant*> incl;
  NodePool<Variant*> excl;
  NodePool<int> sync;
};

struct HalfPath {
  int H = 0;
  HaplotypePlayback hap[MAXH];
  Node<OrientedVariant*>* included = nullptr;
  Node<Variant*>* excluded = nullptr;
  int variantIndex = -1;             // :55
  int variantEndPosition = 0;        // :56
  int includedVariantEndPosition = 0;// :57

  bool isNew(const OrientedVariant& var) const {  // :140-151
    if (var.v->start >= includedVariantEndPosition) return true;
    for (int i = 0; i < H; ++i) {
      if (!hap[i].isNew(var.allele(i))) return false;
    }
    return true;
  }
  void exclude(Pools& P, Variant* var, int varIndex) {  // :153-158
    Node<Variant*>* n = P.excl.push(var, excluded);
    P.excl.release(excluded);
    excluded = n;
    variantEndPosition = std::max(variantEndPosition, var->end);
    variantIndex = varIndex;
  }
  void include(Pools& P, OrientedVariant* var, int varIndex) {  // :160-170
    Node<OrientedVariant*>* n = P.incl.push(var, included);
    P.incl.release(included);
    included = n;
    variantEndPosition = std::max(variantEndPosition, var->v->end);
    variantIndex = varIndex;
    includedVariantEndPosition = std::max(includedVariantEndPosition, var->v->end);
    for (int i = 0; i < H; ++i) hap[i].add(var->allele(i));
  }
  bool wantsFutureVariantBases() const {  // :176-178
    for (int i = 0; i < H; ++i) if (hap[i].wantsFutureVariantBases()) return true;
    return false;
  }
  void step() { for (int i = 0; i < H; ++i) hap[i].step(); }  // :183-187
  void step(int h) { hap[h].step(); }                          // :193-195
  bool finished() const {                                       // :200-202
    for (int i = 0; i < H; ++i) if (!hap[i].finished) return false;
    return true;
  }
  int getPosition() const {  // :256-262
    int m = hap[0].templatePosition;
    for (int i = 1; i < H; ++i) m = std::max(m, hap[i].templatePosition);
    return m;
  }
  void moveForward(int position) { for (int i = 0; i < H; ++i) hap[i].moveForward(position); }  // :272-276
  bool matches(const HalfPath& o) const {  // :278-286
    for (int i = 0; i < H; ++i) if (!hap[i].matches(o.hap[i])) return false;
    return true;
  }
  int compareTo(const HalfPath& that) const {  // :289-298
    for (int i = 0; i < H; ++i) {
      const int plus = hap[i].compareTo(that.hap[i]);
      if (plus != 0) return plus;
    }
    return 0;
  }
  int trailingHaplotype() const {  // :304-318
    int minh = 0;
    int minp = hap[0].templatePosition;
    int maxp = hap[0].templatePosition;
    for (int i = 1; i < H; ++i) {
      const int p = hap[i].templatePosition;
      if (p < minp) { minp = p; minh = i; }
      else if (p > minp) { maxp = p; }
    }
    return minp == maxp ? -1 : minh;
  }
  bool isOnTemplate() const {  // :324-326
    for (int i = 0; i < H; ++i) if (!hap[i].isOnTemplate()) return false;
    return true;
  }
};

// -------------------------------------------------------------------- Path ---
// E/Path.java
struct Path {
  HalfPath called, baseline;
  Node<int>* syncPointList = nullptr;
  int cSinceSync = 0, bSinceSync = 0;

  bool finished() const { return called.finished() && baseline.finished(); }  // :190-192
  bool inSync() const {  // :197-220
    if (called.trailingHaplotype() != -1) return false;
    if (baseline.trailingHaplotype() != -1) return false;
    if (called.getPosition() != baseline.getPosition()) return false;
    if (called.getPosition() < called.includedVariantEndPosition) return false;
    if (baseline.getPosition() < baseline.includedVariantEndPosition) return false;
    if (!called.isOnTemplate()) return false;
    if (!baseline.isOnTemplate()) return false;
    return true;
  }
  bool hasNoOp() const {  // :227-231
    return (bSinceSync == 0 && cSinceSync > 0) || (cSinceSync == 0 && bSinceSync > 0);
  }
  void moveForward(int position) { called.moveForward(position); baseline.moveForward(position); }  // :325-328
  void step() {  // :330-341
    const int trailing = called.trailingHaplotype();
    if (trailing == -1) { called.step(); baseline.step(); }
    else { called.step(trailing); baseline.step(trailing); }
  }
  bool matches() const { return called.matches(baseline); }  // :346-348
  int compareTo(const Path& o) const {  // :351-357
    const int a = called.compareTo(o.called);
    if (a != 0) return a;
    return baseline.compareTo(o.baseline);
  }
};

struct PathLess {
  bool operator()(const Path* a, const Path* b) const { return a->compareTo(*b) < 0; }
};

// ----------------------------------------------------------- PathPreference ---
static int listSize(const Node<OrientedVariant*>* n) { return n ? n->size : 0; }

// E/MaxSumBoth.java:43-82
static Path* betterSum(Path* first, Path* second) {
  int firstSize = listSize(first->called.included);
  int secondSize = listSize(second->called.included);
  const Node<OrientedVariant*>* fi = first->baseline.included;
  const Node<OrientedVariant*>* si = second->baseline.included;
  firstSize += listSize(fi);
  secondSize += listSize(si);
  if (firstSize == secondSize) {
    if (fi != nullptr && si != nullptr) {
      const int fDelta = std::abs(first->bSinceSync - first->cSinceSync);
      const int sDelta = std::abs(second->bSinceSync - second->cSinceSync);
      if (fDelta != sDelta) return fDelta < sDelta ? first : second;
      const int syncDelta = (first->syncPointList ? first->syncPointList->value : 0)
                          - (second->syncPointList ? second->syncPointList->value : 0);
      if (syncDelta != 0) return syncDelta > 0 ? first : second;
      return fi->value->alleleId() < si->value->alleleId() ? first : second;
    }
  }
  return firstSize > secondSize ? first : second;
}

// E/MaxCallsMinBaseline.java:42-61
static Path* betterCallsMinBase(Path* first, Path* second) {
  int firstSize = listSize(first->called.included);
  int secondSize = listSize(second->called.included);
  if (firstSize == secondSize) {
    const Node<OrientedVariant*>* fi = first->baseline.included;
    const Node<OrientedVariant*>* si = second->baseline.included;
    firstSize = listSize(fi);
    secondSize = listSize(si);
    if (firstSize == secondSize) {
      if (fi != nullptr && si != nullptr) {
        return fi->value->alleleId() < si->value->alleleId() ? first : second;
      }
    }
    return firstSize < secondSize ? first : second;
  }
  return firstSize > secondSize ? first : second;
}

// -------------------------------------------------------------- PathFinder ---
// E/PathFinder.java
struct Warning { int paths, iterations, start1, end1; };

struct BestPath {
  bool valid = false;
  std::vector<OrientedVariant*> calledIncluded, baselineIncluded;
  std::vector<Variant*> calledExcluded, baselineExcluded;
  std::vector<int> syncPoints;
};

struct PathFinder {
  const vce_config& cfg;
  int pass;
  const uint8_t* tmpl;
  int tmplLen;
  std::vector<Variant*> calledVariants, baselineVariants;
  int callOrientor, baselineOrientor, H;
  int currentMaxPos = 0;
  Pools P;
  std::vector<Path*> pathFree;
  std::vector<std::unique_ptr<Path>> pathStore;
  std::vector<Warning> warnings;
  long long totalIterations = 0;
  long long syncRegions = 0;

  PathFinder(const vce_config& c, int pass_, const uint8_t* t, int tl,
             const std::vector<Variant*>& base, const std::vector<Variant*>& calls)
      : cfg(c), pass(pass_), tmpl(t), tmplLen(tl), calledVariants(calls), baselineVariants(base) {
    callOrientor = c.calls_orientor[pass_];
    baselineOrientor = c.baseline_orientor[pass_];
    H = c.ploidy[pass_];
    // PathFinder.java:122-126: stable sort by SequenceNameLocusComparator (start, then end)
    auto cmp = [](const Variant* a, const Variant* b) {
      if (a->start != b->start) return a->start < b->start;
      return a->end < b->end;
    };
    std::stable_sort(calledVariants.begin(), calledVariants.end(), cmp);
    std::stable_sort(baselineVariants.begin(), baselineVariants.end(), cmp);
  }

  Path* newPath() {
    if (pathFree.empty()) {
      pathStore.emplace_back(new Path());
      return pathStore.back().get();
    }
    Path* p = pathFree.back();
    pathFree.pop_back();
    return p;
  }
  void freePath(Path* p) {
    P.incl.release(p->called.included);
    P.excl.release(p->called.excluded);
    P.incl.release(p->baseline.included);
    P.excl.release(p->baseline.excluded);
    P.sync.release(p->syncPointList);
    p->called.included = p->baseline.included = nullptr;
    p->called.excluded = p->baseline.excluded = nullptr;
    p->syncPointList = nullptr;
    pathFree.push_back(p);
  }
  // Path(Path parent, syncPoints) E/Path.java:178-185 + HalfPath(HalfPath) :75-85
  Path* copyPath(const Path* parent, Node<int>* syncPoints) {
    Path* p = newPath();
    *p = *parent;
    NodePool<OrientedVariant*>::retain(p->called.included);
    NodePool<Variant*>::retain(p->called.excluded);
    NodePool<OrientedVariant*>::retain(p->baseline.included);
    NodePool<Variant*>::retain(p->baseline.excluded);
    p->syncPointList = NodePool<int>::retain(syncPoints);
    return p;
  }

  Path* better(Path* a, Path* b) const {
    return cfg.preference == VCE_PREFER_CALLS_MIN_BASE ? betterCallsMinBase(a, b) : betterSum(a, b);
  }

  std::vector<OrientedVariant>& orients(Variant* v, bool side) {
    if (!v->orientsReady[pass]) {
      v->orients[pass] = orientations(side ? callOrientor : baselineOrientor, H, v);
      v->orientsReady[pass] = true;
    }
    return v->orients[pass];
  }

  // PathFinder.nextVariant :296-309
  static int nextVariant(const HalfPath& path, const std::vector<Variant*>& variants) {
    const int nextIdx = path.variantIndex + 1;
    if (nextIdx >= (int) variants.size()) return -1;
    const Variant* nextVar = variants[nextIdx];
    if (nextVar->start <= path.getPosition() + 1) return nextIdx;
    if (path.wantsFutureVariantBases() && nextVar->start <= path.variantEndPosition) return nextIdx;
    return -1;
  }
  // :285-292
  int futureVariantPos(const HalfPath& path, const std::vector<Variant*>& variants) const {
    const int nextIdx = path.variantIndex + 1;
    if (nextIdx >= (int) variants.size()) return tmplLen - 1;
    return variants[nextIdx]->start;
  }
  // :271-282
  void skipToNextVariant(Path* head) const {
    const int aNext = futureVariantPos(head->called, calledVariants);
    const int bNext = futureVariantPos(head->baseline, baselineVariants);
    const int lastTemplatePos = tmplLen - 1;
    const int nextPos = std::min(std::min(aNext, bNext), lastTemplatePos) - 1;
    if (nextPos > head->called.getPosition()) head->moveForward(nextPos);
  }
  // :255-264
  void skipVariantsTo(HalfPath& path, const std::vector<Variant*>& variants, int maxPos) const {
    int varIndex = path.variantIndex;
    while (varIndex < (int) variants.size() && (varIndex == -1 || variants[varIndex]->start < maxPos)) ++varIndex;
    --varIndex;
    path.variantIndex = varIndex;
    path.moveForward(std::min(maxPos, tmplLen - 1));
  }
  // :216-229 (flag-alternates unsupported in the oracle: rejected up front)
  bool prune(const Path* head) const { return cfg.prune_no_ops && head->hasNoOp(); }

  typedef std::set<Path*, PathLess> PathSet;

  // :317-348
  void addIfBetter(Path* add, PathSet& sortedPaths) {
    auto it = sortedPaths.find(add);
    if (it != sortedPaths.end()) {
      Path* other = *it;
      Path* best = better(add, other);
      if (best == add) {
        sortedPaths.erase(it);
        sortedPaths.insert(add);
        freePath(other);
      } else {
        freePath(add);
      }
    } else {
      sortedPaths.insert(add);
    }
  }

  // Path.addVariant E/Path.java:288-323 followed by addIfBetter of every child (:244, :311-315)
  void addVariant(Path* head, bool side, Variant* var, int varIndex, PathSet& sortedPaths) {
    Node<int>* syncPoints;
    bool ownSync = false;
    if (head->inSync()) {
      const int syncPos = head->called.getPosition();
      if (head->syncPointList != nullptr && syncPos == head->syncPointList->value) {
        syncPoints = head->syncPointList;
      } else {
        syncPoints = P.sync.push(syncPos, head->syncPointList);
        ownSync = true;
      }
      head->cSinceSync = 0;
      head->bSinceSync = 0;
    } else {
      syncPoints = head->syncPointList;
    }
    std::vector<Path*> children;
    Path* exclude = copyPath(head, syncPoints);
    (side ? exclude->called : exclude->baseline).exclude(P, var, varIndex);
    children.push_back(exclude);
    std::vector<OrientedVariant>& os = orients(var, side);
    for (OrientedVariant& o : os) {
      if ((side ? head->called : head->baseline).isNew(o)) {
        Path* include = copyPath(head, syncPoints);
        if (side) { include->called.include(P, &o, varIndex); ++include->cSinceSync; }
        else { include->baseline.include(P, &o, varIndex); ++include->bSinceSync; }
        children.push_back(include);
      }
    }
    if (ownSync) P.sync.release(syncPoints);
    for (Path* c : children) addIfBetter(c, sortedPaths);
  }

  // :231-247.  NOTE: the parent `head` object is dropped by the caller's `continue`
  // unless it is lastSyncPath (kept alive, see bestPath).
  bool enqueueVariant(PathSet& paths, Path* head, bool side) {
    std::vector<Variant*>& variants = side ? calledVariants : baselineVariants;
    HalfPath& halfPath = side ? head->called : head->baseline;
    const int aVarIndex = nextVariant(halfPath, variants);
    if (aVarIndex != -1) {
      Variant* aVar = variants[aVarIndex];
      currentMaxPos = std::max(currentMaxPos, aVar->start);
      addVariant(head, side, aVar, aVarIndex, paths);
      return true;
    }
    return false;
  }

  // :129-214
  BestPath bestPath() {
    PathSet sortedPaths;
    {
      Path* init = newPath();
      *init = Path();
      init->called.H = init->baseline.H = H;
      for (int i = 0; i < H; ++i) {
        init->called.hap[i] = HaplotypePlayback();
        init->baseline.hap[i] = HaplotypePlayback();
        init->called.hap[i].tmpl = init->baseline.hap[i].tmpl = tmpl;
        init->called.hap[i].tmplLen = init->baseline.hap[i].tmplLen = tmplLen;
      }
      sortedPaths.insert(init);
    }
    Path* best = nullptr;
    int currentIterations = 0;
    currentMaxPos = 0;
    Path* lastSyncPath = nullptr;  // Java keeps the object alive by reference; we own it explicitly
    int lastSyncPos = 0;
    while (!sortedPaths.empty()) {
      ++currentIterations;  // :145 (post-increment; the value tested at :162 is the incremented one)
      ++totalIterations;
      Path* head = *sortedPaths.begin();
      sortedPaths.erase(sortedPaths.begin());
      bool headIsLastSync = false;
      if (sortedPaths.empty()) {  // :151-161
        const int currentSyncPos = head->called.getPosition();
        currentIterations = 0;
        lastSyncPos = currentSyncPos;
        if (lastSyncPath != nullptr && lastSyncPath != head) freePath(lastSyncPath);
        lastSyncPath = head;
        headIsLastSync = true;
      } else if ((int) sortedPaths.size() > cfg.max_paths || currentIterations > cfg.max_iterations) {  // :162-169
        warnings.push_back(Warning{(int) sortedPaths.size(), currentIterations, lastSyncPos + 1, currentMaxPos + 2});
        for (Path* p : sortedPaths) freePath(p);
        sortedPaths.clear();
        currentIterations = 0;
        if (head != lastSyncPath) freePath(head);
        head = lastSyncPath;
        headIsLastSync = true;
        skipVariantsTo(head->called, calledVariants, currentMaxPos + 1);
        skipVariantsTo(head->baseline, baselineVariants, currentMaxPos + 1);
      } else {
        headIsLastSync = (head == lastSyncPath);
      }
      // `head` may alias lastSyncPath: Java semantics are by reference, so mutations of
      // head below are visible through lastSyncPath; we keep exactly that aliasing and only
      // avoid freeing an object that lastSyncPath still refers to.
      if (enqueueVariant(sortedPaths, head, true)) {   // :170
        if (!headIsLastSync) freePath(head);
        continue;
      }
      if (enqueueVariant(sortedPaths, head, false)) {  // :173
        if (!headIsLastSync) freePath(head);
        continue;
      }
      head->step();  // :177
      if (head->inSync()) {  // :182-190
        if (prune(head)) {
          if (!headIsLastSync) freePath(head);
          continue;
        }
        skipToNextVariant(head);
      }
      if (head->matches()) {  // :192
        if (head->finished()) {  // :193-201
          if (prune(head)) {
            if (!headIsLastSync) freePath(head);
            continue;
          }
          Node<int>* sp = P.sync.push(head->called.getPosition(), head->syncPointList);
          Path* finishedPath = copyPath(head, sp);
          P.sync.release(sp);
          if (best == nullptr) {
            best = finishedPath;
          } else {
            Path* b = better(best, finishedPath);
            if (b == best) freePath(finishedPath); else { freePath(best); best = finishedPath; }
          }
          if (!headIsLastSync) freePath(head);
        } else {
          // :203 addIfBetter(head): the same object goes back into the set.  If it is rejected in
          // favour of an equal path it is dropped (unless lastSyncPath still refers to it).
          auto it = sortedPaths.find(head);
          if (it != sortedPaths.end()) {
            Path* other = *it;
            Path* b = better(head, other);
            if (b == head) {
              sortedPaths.erase(it);
              sortedPaths.insert(head);
              if (other != lastSyncPath) freePath(other);
            } else if (!headIsLastSync) {
              freePath(head);
            }
          } else {
            sortedPaths.insert(head);
          }
        }
      } else {
        if (!headIsLastSync) freePath(head);
      }
    }
    BestPath r;
    if (best != nullptr) {
      r.valid = true;
      r.calledIncluded = toReversedList(best->called.included);
      r.calledExcluded = toReversedList(best->called.excluded);
      r.baselineIncluded = toReversedList(best->baseline.included);
      r.baselineExcluded = toReversedList(best->baseline.excluded);
      r.syncPoints = toReversedList(best->syncPointList);
    }
    return r;
  }
};

// ------------------------------------------------------- calculateWeights ---
// E/Path.java:385-453
struct SyncPointRec { int pos; double calledTP, baselineTP, baselineInside; };

static void calculateWeights(const std::vector<int>& syncList, const std::vector<OrientedVariant*>& calledTP,
                             const std::vector<OrientedVariant*>& baselineTP) {
  std::vector<SyncPointRec> sps;
  sps.reserve(syncList.size());
  size_t basePos = 0, callPos = 0;
  for (int loc : syncList) {  // getSyncPointsList :385-408
    int baseLineCount = 0, baseLineInside = 0, calledCount = 0;
    while (basePos < baselineTP.size() && baselineTP[basePos]->v->start <= loc) {
      ++baseLineCount;
      if (!(baselineTP[basePos]->v->status & VCE_STATUS_OUTSIDE_EVAL)) ++baseLineInside;
      ++basePos;
    }
    while (callPos < calledTP.size() && calledTP[callPos]->v->start <= loc) {
      ++calledCount;
      ++callPos;
    }
    sps.push_back(SyncPointRec{loc, (double) calledCount, (double) baseLineCount, (double) baseLineInside});
  }
  size_t it = 0;
  const SyncPointRec* sp = &sps[it++];
  for (OrientedVariant* v : calledTP) {  // :433-452
    while (sp->pos < v->v->start && it < sps.size()) sp = &sps[it++];
    if (sp->baselineTP == 0) v->weight = 0;
    else v->weight = sp->baselineInside / sp->calledTP;
  }
}

// -------------------------------------------------------- PhasingEvaluator ---
// E/PhasingEvaluator.java
struct VariantSummary { const Variant* v; bool included; bool phaseAlt;
  bool isPhased() const { return v->phased; }
  int startPos() const { return v->start; }
  bool phase() const { return phaseAlt; }  // only consulted when isPhased() in the reference
};
struct CallIterator {  // :160-195
  const std::vector<OrientedVariant*>& inc;
  const std::vector<Variant*>& exc;
  size_t i = 0, e = 0;
  CallIterator(const std::vector<OrientedVariant*>& a, const std::vector<Variant*>& b) : inc(a), exc(b) {}
  bool hasNext() const { return i < inc.size() || e < exc.size(); }
  VariantSummary next() {
    const OrientedVariant* ci = i < inc.size() ? inc[i] : nullptr;
    const Variant* ce = e < exc.size() ? exc[e] : nullptr;
    if (ci == nullptr || (ce != nullptr && ce->start < ci->v->start)) {
      ++e;
      return VariantSummary{ce, false, false};
    }
    ++i;
    return VariantSummary{ci->v, true, ci->original};
  }
};

static bool groupInPhase(const std::vector<VariantSummary>& group) {  // :143-155
  if (group.empty()) return true;
  if (!group[0].isPhased()) return false;
  const bool phase = group[0].phase();
  for (const VariantSummary& v : group) {
    if (!v.isPhased() || v.phase() != phase) return false;
  }
  return true;
}

static void countMisphasings(const BestPath& best, int out[3]) {  // :55-142
  CallIterator baseline(best.baselineIncluded, best.baselineExcluded);
  CallIterator calls(best.calledIncluded, best.calledExcluded);
  int misPhasings = 0, unphaseable = 0, correctPhasings = 0;
  bool baseIsPhased = false, basePhase = false, callIsPhased = false, callPhase = false;
  bool haveCall = false, haveBase = false;
  VariantSummary call{nullptr, false, false}, base{nullptr, false, false};
  for (int point : best.syncPoints) {
    std::vector<VariantSummary> baselineSection;
    do {
      if (haveBase && base.startPos() < point) {
        baselineSection.push_back(base);
        haveBase = baseline.hasNext();
        if (haveBase) base = baseline.next();
      }
      if (!haveBase) {
        haveBase = baseline.hasNext();
        if (haveBase) base = baseline.next();
      }
    } while (haveBase && base.startPos() < point);
    std::vector<VariantSummary> callSection;
    do {
      if (haveCall) callSection.push_back(call);
      haveCall = calls.hasNext();
      if (haveCall) call = calls.next();
    } while (haveCall && call.startPos() < point);

    if (!groupInPhase(baselineSection)) {
      for (const VariantSummary& s : callSection) if (s.isPhased()) ++unphaseable;
      baseIsPhased = false;
      callIsPhased = false;
      continue;
    }
    bool transition = false;
    if (!baseIsPhased) callIsPhased = false;
    for (const VariantSummary& bs : baselineSection) {
      if (bs.isPhased()) {
        if (basePhase != bs.phase()) transition = true;
        baseIsPhased = true;
      }
      basePhase = bs.phase();  // groupInPhase guarantees isPhased here (else the reference would throw)
    }
    for (const VariantSummary& cc : callSection) {
      if (!cc.isPhased()) {
        callIsPhased = false;
      } else if (!callIsPhased) {
        callIsPhased = true;
        callPhase = cc.phase();
      } else {
        const bool callTransition = cc.phase() != callPhase;
        if (cc.included && !(callTransition == transition)) {
          ++misPhasings;
          callPhase = cc.phase();
        } else if (cc.included) {
          ++correctPhasings;
          callPhase = cc.phase();
        }
      }
      transition = false;
    }
  }
  out[0] = misPhasings;
  out[1] = correctPhasings;
  out[2] = unphaseable;
}

// ------------------------------------------------------- SequenceEvaluator ---
struct SideData {
  std::vector<Allele> alleleStore;
  std::vector<Variant> variants;
};

static void loadSide(const vce_variants& in, SideData& out) {
  out.alleleStore.resize(in.n_slots);
  for (int s = 0; s < in.n_slots; ++s) {
    Allele& a = out.alleleStore[s];
    a.start = in.allele_start[s];
    a.end = in.allele_end[s];
    a.nt = in.nt + in.allele_nt_off[s];
    a.len = in.allele_nt_off[s + 1] - in.allele_nt_off[s];
  }
  out.variants.resize(in.n);
  for (int i = 0; i < in.n; ++i) {
    Variant& v = out.variants[i];
    v.id = in.id[i];
    v.phased = in.phased ? in.phased[i] != 0 : false;
    v.status = in.status_in ? in.status_in[i] : 0;
    v.inputIndex = i;
    v.gtPloidy = in.gt_ploidy;
    for (int h = 0; h < MAXH; ++h) v.gt[h] = h < in.gt_ploidy ? in.gt[(size_t) i * in.gt_ploidy + h] : -2;
    int start = INT32_MAX, end = INT32_MIN;  // Allele.getAlleleBounds :195-206
    for (int s = in.allele_off[i]; s < in.allele_off[i + 1]; ++s) {
      if (in.allele_null[s]) {
        v.alleles.push_back(nullptr);
      } else {
        v.alleles.push_back(&out.alleleStore[s]);
        start = std::min(start, out.alleleStore[s].start);
        end = std::max(end, out.alleleStore[s].end);
      }
    }
    v.start = start;
    v.end = end;
  }
}

// InterleavingEvalSynchronizer.floorSyncPos :71-83 (stateful walk kept as in the reference)
static int floorSyncPos(const std::vector<int>& syncPoints, int vPos, int sId) {
  const int vv = vPos - 1;
  const int lim = (int) syncPoints.size() - 1;
  int newId = sId;
  while (newId > 0 && syncPoints[newId] > vv) --newId;
  while (newId < lim && syncPoints[newId + 1] <= vv) ++newId;
  return newId;
}

static void fillSide(const SideData& side, const std::vector<OrientedVariant*>& included,
                     const std::vector<OrientedVariant*>& partial, const std::vector<Variant*>& excluded,
                     const std::vector<int>& sync1, const std::vector<int>& sync2, vce_side_result& out) {
  // SequenceEvaluator.mergeVariants :170-182 + insertSkipped :190-209
  const int n = (int) side.variants.size();
  std::vector<uint8_t> seen(n, 0);
  std::vector<const OrientedVariant*> chosen(n, nullptr);
  std::vector<Variant>& vars = const_cast<std::vector<Variant>&>(side.variants);
  for (OrientedVariant* o : included) { o->v->status |= VCE_STATUS_GT_MATCH; seen[o->v->inputIndex] = 1; chosen[o->v->inputIndex] = o; }
  for (OrientedVariant* o : partial) { o->v->status |= VCE_STATUS_ALLELE_MATCH; seen[o->v->inputIndex] = 1; chosen[o->v->inputIndex] = o; }
  for (Variant* v : excluded) { v->status |= VCE_STATUS_NO_MATCH; seen[v->inputIndex] = 1; }
  int sid1 = 0, sid2 = 0;
  for (int i = 0; i < n; ++i) {
    Variant& v = vars[i];
    if (!seen[i]) v.status |= VCE_STATUS_SKIPPED;
    out.status[i] = v.status;
    for (int h = 0; h < MAXH; ++h) out.allele_ids[(size_t) i * MAXH + h] = -2;
    out.original[i] = 0;
    out.weight[i] = 0;
    if (chosen[i]) {
      for (int h = 0; h < chosen[i]->nIds; ++h) out.allele_ids[(size_t) i * MAXH + h] = (int8_t) chosen[i]->alleleIds[h];
      out.original[i] = chosen[i]->original ? 1 : 0;
      out.weight[i] = chosen[i]->weight;
    }
    if (out.sync_id) {
      // the SYNC annotation WithInfoEvalSynchronizer would print (:121-222) = syncStart + 1, or 0 when none
      int s1 = 0, s2 = 0;
      if (!sync1.empty()) { sid1 = floorSyncPos(sync1, v.start, sid1); s1 = sync1[sid1] + 1; }
      if (!sync2.empty()) { sid2 = floorSyncPos(sync2, v.start, sid2); s2 = sync2[sid2] + 1; }
      int val = 0;
      if (v.status & VCE_STATUS_SKIPPED) val = 0;
      else if (v.status & VCE_STATUS_GT_MATCH) val = s1 + 1;
      else if (v.status & VCE_STATUS_ALLELE_MATCH) val = s2 + 1;
      else if (v.status & VCE_STATUS_NO_MATCH) val = s2 > 0 ? s2 + 1 : s1 + 1;
      out.sync_id[i] = val;
    }
  }
}

static int writeSync(const std::vector<int>& sp, vce_sequence_result& res, int pass) {
  res.n_sync[pass] = (int) sp.size();
  if ((int) sp.size() > res.sync_cap[pass]) return VCE_ERR_CAPACITY;
  for (size_t i = 0; i < sp.size(); ++i) res.sync_points[pass][i] = sp[i];
  return VCE_OK;
}

// E/SequenceEvaluator.java:76-158
static int evaluateSequence(const vce_config& cfg, const vce_sequence& seq, vce_sequence_result& res) {
  res.n_sync[0] = res.n_sync[1] = 0;
  res.phasing[0] = res.phasing[1] = res.phasing[2] = 0;
  res.n_warnings = 0;
  res.error = VCE_OK;
  res.n_regions = 0;
  res.n_iterations = 0;
  if (cfg.flag_alternates) { res.error = VCE_ERR_UNSUPPORTED; return res.error; }
  SideData base, calls;
  loadSide(seq.baseline, base);
  loadSide(seq.calls, calls);
  std::vector<int> sync1, sync2;
  std::vector<OrientedVariant*> none;
  if (base.variants.empty() || calls.variants.empty()) {  // :94-97 short circuit
    std::vector<Variant*> be, ce;
    for (Variant& v : base.variants) be.push_back(&v);
    for (Variant& v : calls.variants) ce.push_back(&v);
    fillSide(base, none, none, be, sync1, sync2, res.baseline);
    fillSide(calls, none, none, ce, sync1, sync2, res.calls);
    return VCE_OK;
  }
  try {
    std::vector<Variant*> bv, cv;
    for (Variant& v : base.variants) bv.push_back(&v);
    for (Variant& v : calls.variants) cv.push_back(&v);
    PathFinder f(cfg, 0, seq.template_bases, seq.template_len, bv, cv);
    BestPath best = f.bestPath();
    auto pushWarnings = [&](const PathFinder& pf, int pass) {
      for (const Warning& w : pf.warnings) {
        if (res.n_warnings < res.warn_cap) res.warnings[res.n_warnings] = vce_warning{pass, w.paths, w.iterations, w.start1, w.end1};
        ++res.n_warnings;
      }
      res.n_iterations += pf.totalIterations;
    };
    pushWarnings(f, 0);
    if (!best.valid) { res.error = VCE_ERR_BEST_PATH_NULL; return res.error; }  // :105-116
    countMisphasings(best, res.phasing);  // :122-123
    calculateWeights(best.syncPoints, best.calledIncluded, best.baselineIncluded);  // :127
    std::vector<Variant*> falsePositives = best.calledExcluded;
    std::vector<Variant*> falseNegatives = best.baselineExcluded;
    std::vector<OrientedVariant*> halfPositives, baselineHalfPositives;
    sync1 = best.syncPoints;
    res.n_regions = (int64_t) sync1.size();
    if (cfg.n_passes == 2) {  // :135-149
      PathFinder f2(cfg, 1, seq.template_bases, seq.template_len, falseNegatives, falsePositives);
      BestPath bestHap = f2.bestPath();
      pushWarnings(f2, 1);
      if (!bestHap.valid) { res.error = VCE_ERR_BEST_PATH_NULL; return res.error; }
      halfPositives = bestHap.calledIncluded;
      baselineHalfPositives = bestHap.baselineIncluded;
      calculateWeights(bestHap.syncPoints, halfPositives, baselineHalfPositives);
      falsePositives = bestHap.calledExcluded;
      falseNegatives = bestHap.baselineExcluded;
      sync2 = bestHap.syncPoints;
      fillSide(base, best.baselineIncluded, baselineHalfPositives, falseNegatives, sync1, sync2, res.baseline);
      fillSide(calls, best.calledIncluded, halfPositives, falsePositives, sync1, sync2, res.calls);
    } else {
      fillSide(base, best.baselineIncluded, baselineHalfPositives, falseNegatives, sync1, sync2, res.baseline);
      fillSide(calls, best.calledIncluded, halfPositives, falsePositives, sync1, sync2, res.calls);
    }
    int rc = writeSync(sync1, res, 0);
    if (rc == VCE_OK) rc = writeSync(sync2, res, 1);
    if (rc != VCE_OK) { res.error = rc; return rc; }
  } catch (const OracleError& e) {
    res.error = e.code;
    return e.code;
  }
  return VCE_OK;
}

// ------------------------------------------------------------ RocContainer ---
// E/RocContainer.java:225-318,438-464 ; E/RocPoint.java:70-74
struct RocPoint { double threshold, tp, fp, tpRaw; };
struct DescCmp {  // DescendingDoubleComparator :447-457 (Double.compareTo semantics: -0.0 < 0.0)
  static int dcmp(double a, double b) {
    if (a < b) return -1;
    if (a > b) return 1;
    int64_t ab, bb;
    std::memcpy(&ab, &a, 8);
    std::memcpy(&bb, &b, 8);
    const bool an = std::isnan(a), bn = std::isnan(b);
    if (an || bn) return an == bn ? 0 : (an ? 1 : -1);
    return ab == bb ? 0 : (ab < bb ? -1 : 1);
  }
  bool operator()(double o1, double o2) const {
    if (std::isnan(o1)) return false;      // NaN is never before anything
    if (std::isnan(o2)) return true;       // everything non-NaN is before NaN
    return dcmp(o2, o1) < 0;
  }
};
struct AscCmp {
  bool operator()(double o1, double o2) const { return DescCmp::dcmp(o1, o2) < 0; }
};

template <typename Cmp>
static int rocRun(const vce_roc_in& in, vce_roc_out& out) {
  std::vector<std::map<double, RocPoint, Cmp>> rocs(in.n_filters);
  int64_t noScore = 0;
  for (int64_t i = 0; i < in.n; ++i) {
    double score = in.score[i];
    if (std::isnan(score) || std::isinf(score)) {  // :232-237
      ++noScore;
      score = std::nan("");
    }
    const uint32_t mask = in.filter_mask ? in.filter_mask[i] : 0xffffffffu;
    for (int f = 0; f < in.n_filters; ++f) {
      if (!(mask >> f & 1)) continue;
      auto it = rocs[f].find(score);  // :251-258
      if (it != rocs[f].end()) {
        it->second.tp += in.tp[i];   // RocPoint.add :70-74
        it->second.fp += in.fp[i];
        it->second.tpRaw += in.tp_raw[i];
      } else {
        rocs[f].emplace(score, RocPoint{score, in.tp[i], in.fp[i], in.tp_raw[i]});
      }
    }
  }
  *out.no_score = noScore;
  int rc = VCE_OK;
  for (int f = 0; f < in.n_filters; ++f) {
    RocPoint cum{0, 0, 0, 0};
    int64_t r = 0;
    for (const auto& me : rocs[f]) {  // :296-309
      cum.tp += me.second.tp;
      cum.fp += me.second.fp;
      cum.tpRaw += me.second.tpRaw;
      if (r < out.cap) {
        const size_t o = (size_t) f * out.cap + r;
        out.threshold[o] = me.second.threshold;
        out.cum_tp[o] = cum.tp;
        out.cum_fp[o] = cum.fp;
        out.cum_tp_raw[o] = cum.tpRaw;
      } else {
        rc = VCE_ERR_CAPACITY;
      }
      ++r;
    }
    out.n_rows[f] = r;
  }
  return rc;
}

}  // namespace

// ================================================================ C entry ===
extern "C" {

int oracle_vce_submit(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_sequence_result* results) {
  int rc = VCE_OK;
  for (int i = 0; i < n_seq; ++i) {
    const int r = evaluateSequence(*cfg, seqs[i], results[i]);
    if (r != VCE_OK && rc == VCE_OK) rc = r;
  }
  return rc;
}

// One worker per reference sequence on `threads` threads: the granularity of the reference's
// SimpleThreadPool (E/VcfEvalTask.java:131-137).
int oracle_vce_submit_mt(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_sequence_result* results, int32_t threads) {
  if (threads <= 1) return oracle_vce_submit(cfg, n_seq, seqs, results);
  std::atomic<int> next(0);
  std::atomic<int> rc(VCE_OK);
  // largest sequences first so the pool drains evenly
  std::vector<int> order(n_seq);
  for (int i = 0; i < n_seq; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    return (int64_t) seqs[a].baseline.n + seqs[a].calls.n > (int64_t) seqs[b].baseline.n + seqs[b].calls.n;
  });
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&]() {
      while (true) {
        const int k = next.fetch_add(1);
        if (k >= n_seq) break;
        const int i = order[k];
        const int r = evaluateSequence(*cfg, seqs[i], results[i]);
        if (r != VCE_OK) rc.store(r);
      }
    });
  }
  for (auto& t : pool) t.join();
  return rc.load();
}

int oracle_vce_roc(const vce_roc_in* in, vce_roc_out* out) {
  return in->ascending ? rocRun<AscCmp>(*in, *out) : rocRun<DescCmp>(*in, *out);
}

// Unit-test hook: Orientor.orientations for one variant described by (gt, per-slot null flags).
// Writes rows of (H allele ids, isOriginal) and returns the row count.  OrientorTest.java:49-232.
int oracle_orientations(int32_t kind, int32_t H, int32_t gt_ploidy, const int8_t* gt, int32_t phased,
                        int32_t n_slots, const uint8_t* slot_null, int8_t* out_ids, uint8_t* out_original, int32_t cap) {
  std::vector<Allele> store(n_slots);
  Variant v;
  v.phased = phased != 0;
  v.gtPloidy = gt_ploidy;
  for (int h = 0; h < MAXH; ++h) v.gt[h] = h < gt_ploidy ? gt[h] : -2;
  for (int s = 0; s < n_slots; ++s) v.alleles.push_back(slot_null[s] ? nullptr : &store[s]);
  std::vector<OrientedVariant> os = orientations(kind, H, &v);
  int n = 0;
  for (const OrientedVariant& o : os) {
    if (n < cap) {
      for (int h = 0; h < MAXH; ++h) out_ids[n * MAXH + h] = h < o.nIds ? (int8_t) o.alleleIds[h] : -2;
      out_original[n] = o.original;
    }
    ++n;
  }
  return n;
}

// Unit-test hook: replay one haplotype (HaplotypePlaybackTest.java:42-54): add the alleles in order,
// then step from the start emitting nt() after every next().  Returns the number of bases, or
// -VCE_ERR_* on a reference exception.
int oracle_replay(const uint8_t* tmpl, int32_t len, int32_t n_alleles, const int32_t* starts, const int32_t* ends,
                  const int32_t* nt_off, const uint8_t* nt, uint8_t* out, int32_t cap) {
  std::vector<Allele> al(n_alleles);
  HaplotypePlayback hp;
  hp.tmpl = tmpl;
  hp.tmplLen = len;
  for (int i = 0; i < n_alleles; ++i) {
    al[i].start = starts[i];
    al[i].end = ends[i];
    al[i].nt = nt + nt_off[i];
    al[i].len = nt_off[i + 1] - nt_off[i];
    hp.add(&al[i]);
  }
  int n = 0;
  try {
    while (hp.hasNext()) {
      hp.next();
      if (n < cap) out[n] = hp.nt();
      ++n;
    }
  } catch (const OracleError& e) {
    return -e.code;
  }
  return n;
}

}  // extern "C"
