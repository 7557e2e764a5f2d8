package uk.ac.ox.well.cortexjdk.utils.io.graph.cortex;

import uk.ac.ox.well.cortexjdk.utils.exceptions.CortexJDKException;
import uk.ac.ox.well.cortexjdk.utils.io.graph.DeBruijnGraph;
import uk.ac.ox.well.cortexjdk.utils.kmer.CanonicalKmer;
import uk.ac.ox.well.cortexjdk.utils.kmer.CortexByteKmer;

import java.io.File;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.Collection;
import java.util.Iterator;
import java.util.List;

/**
 * Drop-in for the reference's uk.ac.ox.well.cortexjdk.utils.io.graph.cortex.CortexGraph (same FQCN, same public
 * surface: DeBruijnGraph.java:16-53 plus getCacheHitsByIndex/ByKmer), backed by libcorticall_cuda: the record array
 * lives in B200 HBM; iteration is fed by the streaming decode kernel in blocks, findRecord by the device search.
 * CortexRecord, CortexHeader, CortexColor, CanonicalKmer, CortexByteKmer stay the reference's own Java classes.
 *
 * Additive batch entry points (what FindROIs / Call should call): findRecordIndices, findWindows,
 * containsWindows, findNovel, writeRois.
 *
 * Differences from the reference, all documented in DESIGN.md: no LRU (cache-hit counters stay 0); an unsorted
 * graph is rejected on the first lookup instead of only when a probe happens to notice; graphs with N <= 2
 * records return the exact match (the reference's search loop never runs for them); findRecord issued inside an
 * iteration does not disturb the iterator (the reference's does, CortexGraph.java:172-181,283-289).
 */
public class CortexGraph implements DeBruijnGraph {
    private static final int BLOCK = 1 << 16;      // records decoded per native call while iterating

    private final File cortexFile;
    private final long handle;
    private final CortexHeader header = new CortexHeader();
    private final long numRecords, dataOffset, recordSize;

    private long recordsSeen = 0;
    private CortexRecord nextRecord = null;

    // the decoded block the iterator is in
    private long blockFirst = -1;
    private int blockCount = 0;
    private long[] blockKmers;
    private int[] blockCov;
    private byte[] blockEdges;

    public CortexGraph(String cortexFilePath) { this(new File(cortexFilePath)); }

    public CortexGraph(File cortexFile) {
        this(cortexFile, NativeCortex.open(cortexFile.getAbsolutePath(), Integer.getInteger("corticall.cuda.device", 0)));
    }

    /** Wraps a device-resident graph the library returned (join, sort, the pre-filters); cortexFile is null until it is written. */
    private CortexGraph(File cortexFile, long nativeHandle) {
        this.cortexFile = cortexFile;
        this.handle = nativeHandle;
        long[] h = NativeCortex.header(handle);
        header.setVersion((int) h[0]);
        header.setKmerSize((int) h[1]);
        header.setKmerBits((int) h[2]);
        header.setNumColors((int) h[3]);
        numRecords = h[4];
        dataOffset = h[5];
        recordSize = h[6];
        for (int c = 0; c < header.getNumColors(); c++) {
            long[] ci = NativeCortex.colorInfo(handle, c);
            CortexColor cc = new CortexColor();
            cc.setSampleName(NativeCortex.colorName(handle, c));
            cc.setCleanedAgainstGraphName(NativeCortex.colorGraphName(handle, c));
            cc.setMeanReadLength((int) ci[0]);
            cc.setTotalSequence(ci[1]);
            cc.setTipClippingApplied(ci[2] != 0);
            cc.setLowCovgSupernodesRemoved(ci[3] != 0);
            cc.setLowCovgKmersRemoved(ci[4] != 0);
            cc.setCleanedAgainstGraph(ci[5] != 0);
            cc.setLowCovSupernodesThreshold((int) ci[6]);
            cc.setLowCovKmerThreshold((int) ci[7]);
            header.addColor(cc);
        }
        position(0);
    }

    // ------------------------------------------------------------------------------------------ seek / iterate
    public long position() { return recordsSeen; }

    public void position(long i) {
        if (i < 0) {
            throw new CortexJDKException("Record index is prefix of range (" + i + " vs 0-" + (numRecords - 1) + ")");
        }
        recordsSeen = i;
        nextRecord = getNextRecord();
    }

    public CortexRecord getRecord(long i) {
        position(i);
        return nextRecord;
    }

    private CortexRecord recordAt(long i) {
        if (i >= numRecords) { return null; }
        if (blockFirst < 0 || i < blockFirst || i >= blockFirst + blockCount) {
            int s = header.getKmerBits(), c = header.getNumColors();
            blockCount = (int) Math.min(BLOCK, numRecords - i);
            if (blockKmers == null || blockKmers.length < blockCount * s) {
                blockKmers = new long[BLOCK * s];
                blockCov = new int[BLOCK * c];
                blockEdges = new byte[BLOCK * c];
            }
            NativeCortex.decodeRecords(handle, i, blockCount, blockKmers, blockCov, blockEdges);
            blockFirst = i;
        }
        int j = (int) (i - blockFirst), s = header.getKmerBits(), c = header.getNumColors();
        return new CortexRecord(Arrays.copyOfRange(blockKmers, j * s, (j + 1) * s), Arrays.copyOfRange(blockCov, j * c, (j + 1) * c),
                                Arrays.copyOfRange(blockEdges, j * c, (j + 1) * c), header.getKmerSize(), s);
    }

    private CortexRecord getNextRecord() {
        if (recordsSeen < numRecords) {
            CortexRecord cr = recordAt(recordsSeen);
            recordsSeen++;
            return cr;
        }
        return null;
    }

    public Iterator<CortexRecord> iterator() { position(0); return this; }
    public boolean hasNext() { return nextRecord != null; }

    public CortexRecord next() {
        CortexRecord current = nextRecord;
        nextRecord = getNextRecord();
        if (nextRecord == null) { close(); }
        return current;
    }

    public void remove() { throw new UnsupportedOperationException(); }

    /** Like the reference (CortexGraph.java:264-270) close() leaves the graph usable; device memory is released by dispose(). */
    public void close() {}

    public void dispose() { NativeCortex.dispose(handle); }

    // ------------------------------------------------------------------------------------------ random access
    public CortexRecord findRecord(byte[] bk) {
        if (bk.length > header.getKmerSize()) { throw new ArrayIndexOutOfBoundsException(header.getKmerSize()); }
        if (bk.length < header.getKmerSize()) { return null; }      // prefix compare can never be equals()
        long[] idx = new long[1];
        NativeCortex.findAscii(handle, bk, 1, idx);
        if (idx[0] < 0) { return null; }
        int s = header.getKmerBits(), c = header.getNumColors();
        long[] k = new long[s]; int[] cov = new int[c]; byte[] ed = new byte[c];
        NativeCortex.decodeRecords(handle, idx[0], 1, k, cov, ed);
        return new CortexRecord(k, cov, ed, header.getKmerSize(), s);
    }

    public CortexRecord findRecord(CortexByteKmer bk) { return findRecord(bk.getKmer()); }
    public CortexRecord findRecord(CanonicalKmer ck) { return findRecord(ck.getKmerAsBytes()); }
    public CortexRecord findRecord(String sk) { return findRecord(sk.getBytes()); }

    // ------------------------------------------------------------------------------------------ batch entry points (additive)
    /** nq ASCII k-mers, row-major, any orientation -> record index per query (-1 = findRecord would return null). */
    public long[] findRecordIndices(byte[] asciiKmers) {
        int nq = asciiKmers.length / header.getKmerSize();
        long[] out = new long[nq];
        NativeCortex.findAscii(handle, asciiKmers, nq, out);
        return out;
    }

    /** Every k-window of a contig (Call.loadChildWalk :2358-2381) in one call. */
    public long[] findWindows(byte[] sequence) {
        long[] out = new long[Math.max(sequence.length - header.getKmerSize() + 1, 0)];
        NativeCortex.findWindows(handle, sequence, out);
        return out;
    }

    /** rois.contains(new CanonicalKmer(window)) for every window (Call.java:191-197, 2425-2451) when this graph is the ROI graph. */
    public boolean[] containsWindows(byte[] sequence) {
        boolean[] out = new boolean[Math.max(sequence.length - header.getKmerSize() + 1, 0)];
        NativeCortex.containsWindows(handle, sequence, out);
        return out;
    }

    /** FindROIs.execute :31-70 in one call: scans every record, writes the 1-colour ROI graph, returns the novel count. */
    public long writeRois(int childColor, List<Integer> parentColors, File out) {
        int[] p = new int[parentColors.size()];
        for (int i = 0; i < p.length; i++) { p[i] = parentColors.get(i); }
        return NativeCortex.writeRoiFile(handle, childColor, p, out.getAbsolutePath());
    }

    // ------------------------------------------------------------------------------------------ next rows: whole-graph operations
    private static int[] toArray(Collection<Integer> colors) {
        int[] a = new int[colors.size()];
        int i = 0;
        for (int c : colors) { a[i++] = c; }
        return a;
    }

    /** CortexCollection / Join (CortexCollection.java:34-62,245-293; commands/utils/Join.java:23-57): the sorted union, colours concatenated. */
    public static CortexGraph join(List<CortexGraph> graphs) {
        long[] hs = new long[graphs.size()];
        for (int i = 0; i < hs.length; i++) { hs[i] = graphs.get(i).handle; }
        return new CortexGraph(null, NativeCortex.join(hs));
    }

    /** Sort (commands/utils/Sort.java:19-50): the same records in ascending k-mer order. */
    public CortexGraph sorted() { return new CortexGraph(null, NativeCortex.sort(handle)); }

    /** CortexGraphWriter over the whole graph (CortexGraphWriter.java:31-138). */
    public void write(File out) { NativeCortex.writeGraph(handle, out.getAbsolutePath()); }

    /** FindLowCoverage.execute :47-58: the records with coverage(0) below the limit. */
    public CortexGraph findLowCoverage(int minCoverage) { return new CortexGraph(null, NativeCortex.findLowCoverage(handle, minCoverage)); }

    /** FindShared.execute :60-109 with this = the pedigree graph: ROI records with coverage in a colour outside child / parents / ignored. */
    public CortexGraph findShared(CortexGraph roi, int childColor, Collection<Integer> parentColors, Collection<Integer> ignoreColors) {
        return new CortexGraph(null, NativeCortex.findShared(handle, roi.handle, childColor, toArray(parentColors), toArray(ignoreColors)));
    }

    /** RecoverExcludedKmers.execute :49-92 with this = the pedigree graph. */
    public CortexGraph recoverExcludedKmers(CortexGraph dirty, int childColor) {
        return new CortexGraph(null, NativeCortex.recoverExcludedKmers(handle, dirty.handle, childColor));
    }

    /** CovStats.execute :46-66: rows {coverage, count} in ascending coverage. */
    public int[][] covStats(int childColor, Collection<Integer> parentColors) {
        int[] flat = NativeCortex.covStats(handle, childColor, toArray(parentColors));
        int[][] rows = new int[flat.length / 2][2];
        for (int i = 0; i < rows.length; i++) { rows[i][0] = flat[2 * i]; rows[i][1] = flat[2 * i + 1]; }
        return rows;
    }

    // ------------------------------------------------------------------------------------------ header getters
    public File getFile() { return cortexFile; }
    public CortexHeader getHeader() { return header; }
    public int getVersion() { return header.getVersion(); }
    public int getKmerSize() { return header.getKmerSize(); }
    public int getKmerBits() { return header.getKmerBits(); }
    public String getSampleName(int color) { return getColor(color).getSampleName(); }
    public int getNumColors() { return header.getNumColors(); }
    public long getNumRecords() { return numRecords; }
    public List<CortexColor> getColors() { return header.getColors(); }
    public boolean hasColor(int color) { return header.hasColor(color); }
    public CortexColor getColor(int color) { return header.getColor(color); }

    public int getColorForSampleName(String sampleName) {
        int sampleColor = -1, copies = 0;
        for (int c = 0; c < header.getNumColors(); c++) {
            if (header.getColor(c).getSampleName().equalsIgnoreCase(sampleName)) { sampleColor = c; copies++; }
        }
        if (sampleColor == -1) {
            try { sampleColor = Integer.valueOf(sampleName); copies = 1; } catch (NumberFormatException e) { /* not a colour index */ }
        }
        return copies == 1 ? sampleColor : -1;
    }

    public List<Integer> getColorsForSampleNames(Collection<String> sampleNames) {
        List<Integer> colors = new ArrayList<>();
        if (sampleNames != null) { for (String s : sampleNames) { colors.add(getColorForSampleName(s)); } }
        return colors;
    }

    public long getCacheHitsByIndex() { return 0; }
    public long getCacheHitsByKmer() { return 0; }

    public String toString() {
        StringBuilder sb = new StringBuilder();
        sb.append("file: ").append(cortexFile.getAbsolutePath()).append("\n----\n")
          .append("binary version: ").append(getVersion()).append("\nkmer size: ").append(getKmerSize())
          .append("\nbitfields: ").append(getKmerBits()).append("\ncolors: ").append(getNumColors()).append("\n");
        for (int c = 0; c < getNumColors(); c++) {
            CortexColor cc = getColor(c);
            sb.append("-- Color ").append(c).append(" --\n  sample name: '").append(cc.getSampleName()).append("'\n")
              .append("  mean read length: ").append(cc.getMeanReadLength()).append("\n");
        }
        sb.append("----\nkmers: ").append(getNumRecords()).append("\n----\n");
        return sb.toString();
    }
}
