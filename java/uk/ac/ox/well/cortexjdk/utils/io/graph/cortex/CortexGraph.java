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
import java.util.LinkedHashMap;
import java.util.List;
import java.util.Map;

/**
 * Drop-in for the reference's uk.ac.ox.well.cortexjdk.utils.io.graph.cortex.CortexGraph (same FQCN, same public
 * surface: DeBruijnGraph.java:16-53 plus getCacheHitsByIndex/ByKmer), backed by libcorticall_cuda: the record array
 * lives in B200 HBM; iteration is fed by the streaming decode kernel in blocks, findRecord by the device search.
 * CortexRecord, CortexHeader, CortexColor, CanonicalKmer, CortexByteKmer stay the reference's own Java classes.
 *
 * Additive batch entry points (what FindROIs / Call should call): findRecordIndices, findWindows,
 * containsWindows, findNovel, writeRois.
 *
 * The legacy per-record findRecord stays cheap for un-patched callers: the traversal engine asks for a vertex and then for
 * its up to eight neighbours (utils/traversal/TraversalEngine.java:67-252), so a miss in the small result cache fetches the
 * k-mer AND its eight possible neighbours in one native call and one kernel launch (NativeCortex.findRecords); the
 * neighbour lookups that follow are answered from the cache.
 *
 * Several GPUs: -Dcorticall.cuda.devices=0,1,...,7 makes new CortexGraph(file) cut the record array into k-mer-range shards,
 * one per device (cc_open_sharded); the batch entry points and findRecord then route their queries to the owning shard.
 *
 * Differences from the reference, all documented in DESIGN.md: no LRU (cache-hit counters stay 0); an unsorted
 * graph is rejected on the first lookup instead of only when a probe happens to notice; graphs with N <= 2
 * records return the exact match (the reference's search loop never runs for them); findRecord issued inside an
 * iteration does not disturb the iterator (the reference's does, CortexGraph.java:172-181,283-289).
 */
public class CortexGraph implements DeBruijnGraph {
    private static final int BLOCK = 1 << 16;      // records decoded per native call while iterating

    private static final int CACHE_SIZE = 1 << 14;
    private static final Object MISSING = new Object();       // cached "findRecord returns null"

    private final File cortexFile;
    private final long handle;                                // cc_graph (one device), or the shard that answers header queries
    private final long sharded;                               // cc_sharded when the graph spans several devices, else 0
    private final long[] shardHandles, shardFirst;            // per-shard cc_graph handles and first record indices
    private final Map<String, Object> findCache = new LinkedHashMap<String, Object>(CACHE_SIZE, 0.75f, true) {
        protected boolean removeEldestEntry(Map.Entry<String, Object> e) { return size() > CACHE_SIZE; }
    };
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
        this(cortexFile, openNative(cortexFile));
    }

    /** {cc_graph handle, cc_sharded handle or 0}: one device unless -Dcorticall.cuda.devices lists several
     *  (-Dcorticall.cuda.placement=auto|range|replicate: replicas of a graph that fits every device, k-mer ranges otherwise). */
    private static long[] openNative(File f) {
        String devs = System.getProperty("corticall.cuda.devices");
        if (devs != null && devs.contains(",")) {
            String[] parts = devs.split(",");
            int[] ids = new int[parts.length];
            for (int i = 0; i < ids.length; i++) { ids[i] = Integer.parseInt(parts[i].trim()); }
            String place = System.getProperty("corticall.cuda.placement", "auto");
            long sh = NativeCortex.openSharded(f.getAbsolutePath(), ids, place.equals("range") ? 0 : place.equals("replicate") ? 1 : 2);
            return new long[] { NativeCortex.shardedShard(sh, 0)[0], sh };
        }
        return new long[] { NativeCortex.open(f.getAbsolutePath(), Integer.getInteger("corticall.cuda.device", 0)), 0 };
    }

    /** Wraps a device-resident graph the library returned (join, sort, the pre-filters); cortexFile is null until it is written. */
    private CortexGraph(File cortexFile, long nativeHandle) { this(cortexFile, new long[] { nativeHandle, 0 }); }

    private CortexGraph(File cortexFile, long[] handles) {
        this.cortexFile = cortexFile;
        this.handle = handles[0];
        this.sharded = handles[1];
        long[] h = NativeCortex.header(handle);
        header.setVersion((int) h[0]);
        header.setKmerSize((int) h[1]);
        header.setKmerBits((int) h[2]);
        header.setNumColors((int) h[3]);
        if (sharded != 0) {
            long[] info = NativeCortex.shardedInfo(sharded);
            shardHandles = new long[(int) info[0]];
            shardFirst = new long[(int) info[0]];
            for (int r = 0; r < shardHandles.length; r++) {
                long[] sh = NativeCortex.shardedShard(sharded, r);
                shardHandles[r] = sh[0];
                shardFirst[r] = sh[2];
            }
            numRecords = info[1];
        } else {
            shardHandles = new long[] { handle };
            shardFirst = new long[] { 0 };
            numRecords = h[4];
        }
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
            // the block never crosses a shard boundary: records [i, i + blockCount) of the shard that holds record i
            int r = shardOf(i);
            long shardEnd = r + 1 < shardFirst.length ? shardFirst[r + 1] : numRecords;
            blockCount = (int) Math.min(blockCount, shardEnd - i);
            NativeCortex.decodeRecords(shardHandles[r], i - shardFirst[r], blockCount, blockKmers, blockCov, blockEdges);
            blockFirst = i;
        }
        int j = (int) (i - blockFirst), s = header.getKmerBits(), c = header.getNumColors();
        return new CortexRecord(Arrays.copyOfRange(blockKmers, j * s, (j + 1) * s), Arrays.copyOfRange(blockCov, j * c, (j + 1) * c),
                                Arrays.copyOfRange(blockEdges, j * c, (j + 1) * c), header.getKmerSize(), s);
    }

    private int shardOf(long i) {
        int r = 0;
        while (r + 1 < shardFirst.length && shardFirst[r + 1] <= i) { r++; }
        return r;
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

    public void dispose() {
        if (sharded != 0) { NativeCortex.disposeSharded(sharded); } else { NativeCortex.dispose(handle); }
    }

    // ------------------------------------------------------------------------------------------ random access
    public CortexRecord findRecord(byte[] bk) {
        if (bk.length > header.getKmerSize()) { throw new ArrayIndexOutOfBoundsException(header.getKmerSize()); }
        if (bk.length < header.getKmerSize()) { return null; }      // prefix compare can never be equals()
        String key = new String(bk);
        Object hit = findCache.get(key);
        if (hit == null) {
            fetchWithNeighbours(bk);
            hit = findCache.get(key);
        }
        return hit == MISSING ? null : (CortexRecord) hit;
    }

    /** The k-mer and its eight possible neighbours (four successors, four predecessors) in one native call; all nine answers are cached. */
    private void fetchWithNeighbours(byte[] bk) {
        final int k = header.getKmerSize(), s = header.getKmerBits(), c = header.getNumColors(), nq = 9;
        final byte[] bases = { 'A', 'C', 'G', 'T' };
        byte[] batch = new byte[nq * k];
        System.arraycopy(bk, 0, batch, 0, k);
        for (int b = 0; b < 4; b++) {
            System.arraycopy(bk, 1, batch, (1 + b) * k, k - 1);               // successor: bk[1..] + base
            batch[(1 + b) * k + k - 1] = bases[b];
            batch[(5 + b) * k] = bases[b];                                    // predecessor: base + bk[..k-1]
            System.arraycopy(bk, 0, batch, (5 + b) * k + 1, k - 1);
        }
        long[] idx = new long[nq];
        long[] kmers = new long[nq * s];
        int[] cov = new int[nq * c];
        byte[] ed = new byte[nq * c];
        if (sharded != 0) {
            // several devices: the indices come from the routed lookup, the records from the owning shards
            NativeCortex.findAsciiSharded(sharded, batch, nq, idx);
            for (int i = 0; i < nq; i++) {
                if (idx[i] < 0) { continue; }
                int r = shardOf(idx[i]);
                long[] k1 = new long[s]; int[] c1 = new int[c]; byte[] e1 = new byte[c];
                NativeCortex.decodeRecords(shardHandles[r], idx[i] - shardFirst[r], 1, k1, c1, e1);
                System.arraycopy(k1, 0, kmers, i * s, s); System.arraycopy(c1, 0, cov, i * c, c); System.arraycopy(e1, 0, ed, i * c, c);
            }
        } else {
            NativeCortex.findRecords(handle, batch, nq, idx, kmers, cov, ed);
        }
        for (int i = 0; i < nq; i++) {
            String key = new String(batch, i * k, k);
            findCache.put(key, idx[i] < 0 ? MISSING
                                          : new CortexRecord(Arrays.copyOfRange(kmers, i * s, (i + 1) * s), Arrays.copyOfRange(cov, i * c, (i + 1) * c),
                                                             Arrays.copyOfRange(ed, i * c, (i + 1) * c), k, s));
        }
    }

    public CortexRecord findRecord(CortexByteKmer bk) { return findRecord(bk.getKmer()); }
    public CortexRecord findRecord(CanonicalKmer ck) { return findRecord(ck.getKmerAsBytes()); }
    public CortexRecord findRecord(String sk) { return findRecord(sk.getBytes()); }

    // ------------------------------------------------------------------------------------------ batch entry points (additive)
    /** nq ASCII k-mers, row-major, any orientation -> record index per query (-1 = findRecord would return null). */
    public long[] findRecordIndices(byte[] asciiKmers) {
        int nq = asciiKmers.length / header.getKmerSize();
        long[] out = new long[nq];
        if (sharded != 0) { NativeCortex.findAsciiSharded(sharded, asciiKmers, nq, out); } else { NativeCortex.findAscii(handle, asciiKmers, nq, out); }
        return out;
    }

    /** Every k-window of a contig (Call.loadChildWalk :2358-2381) in one call. */
    public long[] findWindows(byte[] sequence) {
        long[] out = new long[Math.max(sequence.length - header.getKmerSize() + 1, 0)];
        if (sharded != 0) { NativeCortex.findWindowsSharded(sharded, sequence, out); } else { NativeCortex.findWindows(handle, sequence, out); }
        return out;
    }

    /** rois.contains(new CanonicalKmer(window)) for every window (Call.java:191-197, 2425-2451) when this graph is the ROI graph. */
    public boolean[] containsWindows(byte[] sequence) {
        boolean[] out = new boolean[Math.max(sequence.length - header.getKmerSize() + 1, 0)];
        if (sharded != 0) {
            long[] idx = findWindows(sequence);
            for (int i = 0; i < out.length; i++) { out[i] = idx[i] >= 0; }
        } else {
            NativeCortex.containsWindows(handle, sequence, out);
        }
        return out;
    }

    /** FindROIs.execute :31-70 in one call: scans every record, writes the 1-colour ROI graph, returns the novel count. */
    public long writeRois(int childColor, List<Integer> parentColors, File out) {
        int[] p = new int[parentColors.size()];
        for (int i = 0; i < p.length; i++) { p[i] = parentColors.get(i); }
        return sharded != 0 ? NativeCortex.writeRoiFileSharded(sharded, childColor, p, out.getAbsolutePath())
                            : NativeCortex.writeRoiFile(handle, childColor, p, out.getAbsolutePath());
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

    /** Remove (commands/utils/Remove.java:30-88) with this = the primary graph: the records without coverage in a secondary graph. */
    public CortexGraph remove(List<CortexGraph> secondaries) {
        long[] hs = new long[secondaries.size()];
        for (int i = 0; i < hs.length; i++) { hs[i] = secondaries.get(i).handle; }
        return new CortexGraph(null, NativeCortex.remove(handle, hs)[0]);
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
        sb.append("file: ").append(cortexFile == null ? "<device-resident>" : cortexFile.getAbsolutePath()).append("\n----\n")
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
