"""The two callers of the hot path, with the reference's names and argument meaning, each reduced to ONE batched
device call per graph / per contig:

  FindROIs                S/commands/discover/roi/FindROIs.java:17-106      (the novelty scan + ROI writer)
  CallHelpers.loadRois / loadChildWalk / getRegions / sectionRois
                          S/commands/discover/call/Call.java:2348-2356, 2358-2381, 2425-2451, 191-197

Only the k-mer work of `Call` is here; Tesserae alignment, bubble / breakpoint calling and VCF writing stay in the
Java host (out of scope).
"""
from __future__ import annotations

import os

import numpy as np

from .._native import CortexJDKException
from .cortex import CortexGraph, CortexRecord, packCanonical
from .kmer import CanonicalKmer


class FindROIs:
    """`FindROIs -g trio.ctx -p mom -p dad -c kid -o rois.ctx`: fields named as the reference's @Argument fields."""

    def __init__(self, GRAPH: CortexGraph, PARENTS: list[str], CHILD: str, out):
        self.GRAPH, self.PARENTS, self.CHILD, self.out = GRAPH, list(PARENTS), CHILD, out
        self.numNovelRecords = 0

    def execute(self) -> int:
        childColor = self.GRAPH.getColorForSampleName(self.CHILD)
        parentColors = self.GRAPH.getColorsForSampleNames(self.PARENTS)
        # a sample name that does not resolve gives colour -1, which the reference then uses as an array index
        # (ArrayIndexOutOfBoundsException); the library reports the same condition as CC_ERR_ARG
        self.numNovelRecords = self.GRAPH.writeRois(childColor, parentColors, os.fspath(self.out))
        return self.numNovelRecords


class CortexCollection:
    """S/utils/io/graph/cortex/CortexCollection.java: several graphs presented as one (colours concatenated, records
    merged by k-mer).  The merge is done once on the device (cc_join); iteration, getRecord and findRecord are then
    those of the merged CortexGraph."""

    def __init__(self, *graphs):
        if len(graphs) == 1 and isinstance(graphs[0], (list, tuple)):
            graphs = tuple(graphs[0])
        self.graphList = list(graphs)
        self.merged = CortexGraph.join(self.graphList)

    def __getattr__(self, name):            # DeBruijnGraph surface: delegate to the merged graph
        return getattr(self.merged, name)

    def __iter__(self):
        return iter(self.merged)

    def getGraph(self, color: int) -> CortexGraph:                      # :65-75
        for g in self.graphList:
            if color < g.getNumColors():
                return g
            color -= g.getNumColors()
        raise IndexError(color)


class Join:
    """`Join -g a.ctx -g b.ctx ... -o joined.ctx` (S/commands/utils/Join.java:16-58)."""

    def __init__(self, GRAPHS, out):
        self.GRAPHS, self.out = list(GRAPHS), out

    def execute(self) -> int:
        cc = CortexCollection(self.GRAPHS)
        cc.merged.writeGraph(self.out)
        n = cc.merged.getNumRecords()
        cc.merged.dispose()
        return n


class Remove:
    """`Remove -g primary.ctx -s secondary.ctx ... -o out.ctx` (S/commands/utils/Remove.java:19-88): the primary graph minus
    the k-mers with coverage in a secondary graph."""

    def __init__(self, PGRAPH: CortexGraph, SGRAPH, out):
        self.PGRAPH, self.SGRAPH, self.out = PGRAPH, list(SGRAPH), out

    def execute(self) -> tuple[int, int]:
        """-> (numKept, numRemoved), the two counters the reference logs."""
        kept, removed = self.PGRAPH.remove(self.SGRAPH)
        kept.writeGraph(self.out)
        n = kept.getNumRecords()
        kept.dispose()
        return n, removed


class Sort:
    """`Sort -cg raw.ctx -o sorted.ctx` (S/commands/utils/Sort.java:12-51)."""

    def __init__(self, CORTEX_GRAPH: CortexGraph, out):
        self.CORTEX_GRAPH, self.out = CORTEX_GRAPH, out

    def execute(self) -> int:
        sg = self.CORTEX_GRAPH.sorted()
        sg.writeGraph(self.out)
        n = sg.getNumRecords()
        sg.dispose()
        return n


class FindLowCoverage:
    """`FindLowCoverage -r roi.ctx -m 10 -o low.ctx` (S/commands/prefilter/FindLowCoverage.java:18-67): writes the records
    BELOW the limit (the excluded ones), as the reference does."""

    def __init__(self, ROI: CortexGraph, out, MIN_COVERAGE: int = 10):
        self.ROI, self.out, self.MIN_COVERAGE = ROI, out, MIN_COVERAGE

    def execute(self) -> tuple[int, int]:
        low = self.ROI.findLowCoverage(self.MIN_COVERAGE)
        low.writeGraph(self.out)
        excluded = low.getNumRecords()
        low.dispose()
        return self.ROI.getNumRecords() - excluded, excluded            # (numKept, numExcluded) of the log line


class FindShared:
    """`FindShared -g pedigree.ctx -p mom -p dad -i ref -r roi.ctx -o shared.ctx` (S/commands/prefilter/FindShared.java:23-119)."""

    def __init__(self, GRAPH: CortexGraph, PARENTS: list[str], IGNORE: list[str], ROI: CortexGraph, out):
        self.GRAPH, self.PARENTS, self.IGNORE, self.ROI, self.out = GRAPH, list(PARENTS), list(IGNORE), ROI, out

    def execute(self) -> tuple[int, int]:
        child = self.GRAPH.getColorForSampleName(self.ROI.getSampleName(0))          # :42-43 (may be -1: then nothing is excluded for it)
        parents = self.GRAPH.getColorsForSampleNames(self.PARENTS)
        ignore = self.GRAPH.getColorsForSampleNames(self.IGNORE)
        shared = self.GRAPH.findShared(self.ROI, child, parents, ignore)
        shared.writeGraph(self.out)
        excluded = shared.getNumRecords()
        shared.dispose()
        return self.ROI.getNumRecords() - excluded, excluded


class RecoverExcludedKmers:
    """`RecoverExcludedKmers -g pedigree.ctx -d dirty.ctx -o recovered.ctx` (S/commands/discover/recover/RecoverExcludedKmers.java:17-107)."""

    def __init__(self, GRAPH: CortexGraph, DIRTY: CortexGraph, out):
        self.GRAPH, self.DIRTY, self.out = GRAPH, DIRTY, out

    def execute(self) -> int:
        name = self.DIRTY.getSampleName(0)
        child = self.GRAPH.getColorForSampleName(name)
        if child < 0:                                                                # :33-36
            raise CortexJDKException("Sample '%s' not found in pedigree graph" % name)
        rec, recovered = self.GRAPH.recoverExcludedKmers(self.DIRTY, child)
        rec.writeGraph(self.out)
        rec.dispose()
        return recovered


class CovStats:
    """`CovStats -g pedigree.ctx -c kid -p mom -p dad -o stats.txt` (S/commands/utils/CovStats.java:14-87)."""

    def __init__(self, GRAPH: CortexGraph, CHILD: str, PARENTS, out):
        self.GRAPH, self.CHILD, self.PARENTS, self.out = GRAPH, CHILD, list(PARENTS), out

    def execute(self) -> list[tuple[int, int]]:
        child = self.GRAPH.getColorForSampleName(self.CHILD)
        parents = {self.GRAPH.getColorForSampleName(p) for p in self.PARENTS}        # :78-86
        rows = self.GRAPH.covStats(child, sorted(parents))
        text = "".join("%d\t%d\n" % r for r in rows)                                 # :68-70
        if hasattr(self.out, "write"):
            self.out.write(text)
        else:
            with open(self.out, "w") as f:
                f.write(text)
        return rows


class CortexVertex:
    """The three fields of utils/traversal/CortexVertex the child walk fills (bases, record, copy index)."""

    __slots__ = ("bases", "record", "copyIndex", "recordIndex")

    def __init__(self, bases: str, record, copyIndex: int, recordIndex: int):
        self.bases, self.record, self.copyIndex, self.recordIndex = bases, record, copyIndex, recordIndex

    def getKmerAsString(self): return self.bases
    def getCortexRecord(self): return self.record
    def getCopyIndex(self): return self.copyIndex
    def getCanonicalKmer(self): return CanonicalKmer(self.bases)


class CallHelpers:
    @staticmethod
    def loadRois(rg: CortexGraph) -> CortexGraph:
        """Call.loadRois builds a HashSet<CanonicalKmer> by iterating the ROI graph.  Here the device-resident ROI
        graph IS the set: membership of every window of a contig is one `containsWindows` call."""
        rg.buildIndex()
        return rg

    @staticmethod
    def loadChildWalk(contig: str, graph: CortexGraph, materialize: bool = True) -> list[CortexVertex]:
        """Call.loadChildWalk :2358-2381: one findRecord per window of the contig -> one batched lookup."""
        k = graph.getKmerSize()
        idx = graph.findWindows(contig)
        walk: list[CortexVertex] = []
        seen: dict[str, int] = {}
        recs = {}
        if materialize:
            hits = np.unique(idx[idx >= 0])
            for i in hits.tolist():                      # decode each distinct hit once (device decode)
                w, c, e = graph.decodeRecords(i - graph.firstIndex, 1)
                recs[i] = graph._make_record(w[0], c[0], e[0])
        for i in range(len(idx)):
            sk = contig[i:i + k]
            seen[sk] = seen[sk] + 1 if sk in seen else 0
            walk.append(CortexVertex(sk, recs.get(int(idx[i])), seen[sk], int(idx[i])))
        return walk

    @staticmethod
    def getRegions(rois: CortexGraph, contig: str) -> list[tuple[int, int]]:
        """Call.getRegions :2425-2451: maximal runs of consecutive windows whose canonical k-mer is in the ROI set."""
        present = rois.containsWindows(contig)
        regions, start, stop = [], -1, 0
        for i, p in enumerate(present.tolist()):
            if p:
                if start == -1:
                    start = i
                stop = i
            elif start > -1:
                regions.append((start, stop))
                start, stop = -1, 0
        if start > -1:
            regions.append((start, stop))
        return regions

    @staticmethod
    def sectionRois(rois: CortexGraph, trimmedQuery: str) -> list[CanonicalKmer]:
        """Call.java:191-197: the sorted set of ROI k-mers a query section contains."""
        k = rois.getKmerSize()
        present = rois.containsWindows(trimmedQuery)
        return sorted({CanonicalKmer(trimmedQuery[i:i + k]) for i in np.nonzero(present)[0].tolist()})

    @staticmethod
    def trimQuery(walkContig: str, targets, rois: CortexGraph):
        """Call.trimQuery :1946-1986 on the contig of the child walk `ws` (ws[i] = window i of walkContig): the span from the first
        to the last window that is either novel (its canonical k-mer is in the ROI set: one `containsWindows` call) or shared with
        one of the target sequences (the canonical packed k-mers of all windows come from K3, `packCanonical`; the intersection is
        a set operation on the packed words).  Returns (firstIndex, lastIndex + 1, contig of ws[firstIndex..lastIndex])."""
        k = rois.getKmerSize()
        nw = max(len(walkContig) - k + 1, 0)
        novel = np.nonzero(rois.containsWindows(walkContig))[0]
        firstNovel, lastNovel = (int(novel[0]), int(novel[-1])) if novel.size else (-1, -1)
        firstIndex, lastIndex = 2 ** 31 - 1, 0                                       # Integer.MAX_VALUE, 0
        ww, wf = packCanonical(walkContig, k, rois._device)
        void = np.dtype((np.void, 8 * ww.shape[1]))
        clean = (wf & 6) == 0                                # upper-case ACGT only: equality of CanonicalKmers = equality of packed words
        wkeys = np.ascontiguousarray(ww).view(void).reshape(-1)
        dirty_walk = {}                                      # windows with N / lower case compare as the byte strings they are (rare)
        for i in np.nonzero(~clean)[0].tolist():
            dirty_walk.setdefault(CanonicalKmer(walkContig[i:i + k]).getKmerAsString(), []).append(i)
        for target in (targets.values() if hasattr(targets, "values") else targets):
            tw, tf = packCanonical(target, k, rois._device)
            tclean = (tf & 6) == 0
            tkeys = np.ascontiguousarray(tw[tclean]).view(void).reshape(-1)
            shared = np.nonzero(clean & np.isin(wkeys, tkeys))[0].tolist()
            if dirty_walk:
                for j in np.nonzero(~tclean)[0].tolist():
                    shared += dirty_walk.get(CanonicalKmer(target[j:j + k]).getKmerAsString(), [])
            if shared:
                firstIndex, lastIndex = min(firstIndex, min(shared)), max(lastIndex, max(shared))
        if firstNovel < firstIndex:
            firstIndex = firstNovel
        if lastNovel > lastIndex:
            lastIndex = lastNovel
        if firstIndex < 0 or nw == 0:                        # ws.subList(-1, ...) throws in the reference
            raise IndexError("trimQuery: the walk has neither a novel k-mer nor a k-mer of a target")
        return firstIndex, lastIndex + 1, walkContig[firstIndex:lastIndex + k]

    @staticmethod
    def noveltyMask(rois: CortexGraph, query: str) -> str:
        """The first loop of Call.makeNoveltyTrack :2062-2070: a track one character longer than the (gap-free) query with '*' over
        every base covered by a window whose canonical k-mer is in the ROI set; the gap insertion and expansion that follow in the
        reference work on the alignment columns and stay in the host."""
        k = rois.getKmerSize()
        present = rois.containsWindows(query)
        cover = np.zeros(len(query) + 2, dtype=np.int32)
        hit = np.nonzero(present)[0]
        np.add.at(cover, hit, 1)
        np.add.at(cover, hit + k, -1)
        starred = np.cumsum(cover)[:len(query) + 1] > 0
        return "".join("*" if s else " " for s in starred.tolist())
