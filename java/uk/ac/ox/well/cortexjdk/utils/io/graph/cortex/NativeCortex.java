package uk.ac.ox.well.cortexjdk.utils.io.graph.cortex;

import java.io.File;
import java.io.IOException;
import java.io.InputStream;
import java.nio.file.Files;
import java.nio.file.StandardCopyOption;

/**
 * JNI face of libcorticall_cuda (include/corticall_cuda.h).  One static native method per C-ABI entry point the
 * Java classes need; each forwards to exactly one cc_* function (corticall_b200/csrc/jni_shim.cpp) and throws
 * CortexJDKException carrying cc_last_error() when the status is not CC_OK.
 *
 * Loading follows the reference's own precedent for libbwajni (utils/alignment/pairwise/BwaAligner.java:19-27,
 * utils/packageutils/InternalLibraryResource.java:19-62): the .so travels inside the jar (build.xml already packs
 * **&#47;*.so) and is unpacked to a temporary file; System.load(absolutePath) avoids the usr_paths reflection hack.
 *
 * NOT COMPILED IN THE BUILD CONTAINER: the image has no JDK and no jni.h (SURVEY.md section 0).  The shim is kept
 * to argument marshalling only; every cc_* function it calls is exercised by the Python ctypes tests.
 */
public final class NativeCortex {
    private NativeCortex() {}

    static {
        String explicit = System.getProperty("corticall.cuda.lib");
        try {
            if (explicit != null) {
                System.load(new File(explicit).getAbsolutePath());
            } else {
                try (InputStream in = NativeCortex.class.getResourceAsStream("/libcorticall_cuda.so")) {
                    if (in == null) { throw new IOException("libcorticall_cuda.so is not on the classpath"); }
                    File tmp = File.createTempFile("libcorticall_cuda", ".so");
                    tmp.deleteOnExit();
                    Files.copy(in, tmp.toPath(), StandardCopyOption.REPLACE_EXISTING);
                    System.load(tmp.getAbsolutePath());
                }
            }
        } catch (IOException e) {
            throw new UnsatisfiedLinkError("cannot load libcorticall_cuda: " + e);
        }
    }

    // lifecycle -------------------------------------------------------------------------------- cc_open / cc_dispose
    static native long open(String path, int device);
    static native void dispose(long handle);

    // header ----------------------------------------------------------------------------------- cc_header / cc_color_*
    /** {version, kmerSize, kmerBits, numColors, numRecords, dataOffset, recordSize} */
    static native long[] header(long handle);
    static native String colorName(long handle, int color);
    static native String colorGraphName(long handle, int color);
    /** {meanReadLength, totalSequence, tipClipping, lowCovgSupernodesRemoved, lowCovgKmersRemoved, cleanedAgainstGraph,
     *   lowCovSupernodesThreshold, lowCovKmerThreshold} */
    static native long[] colorInfo(long handle, int color);

    // K1 --------------------------------------------------------------------------------------- cc_decode_records
    /** Decodes records [first, first+count): binaryKmers receives count*kmerBits longs ALREADY byte-swapped to the
     *  Java convention (Long.reverseBytes of the native word, CortexGraph.java:208-209), coverages count*numColors
     *  ints, edges count*numColors bytes. */
    static native void decodeRecords(long handle, long first, int count, long[] binaryKmers, int[] coverages, byte[] edges);

    // K1+K2 ------------------------------------------------------------------------------------ cc_find_novel / cc_write_roi_file
    /** Returns the number of novel records; fills outRecords (8*kmerBits+5 bytes each, writer layout) and outIndex
     *  (may be null) up to their capacity. */
    static native long findNovel(long handle, int child, int[] parents, byte[] outRecords, long[] outIndex);
    static native long writeRoiFile(long handle, int child, int[] parents, String outPath);

    // legacy findRecord, batched ------------------------------------------------------------- cc_find_records
    /** nq k-byte ASCII k-mers (a vertex and its neighbours) in ONE native call and one kernel launch: outIndex[i] = record index
     *  or -1; for every hit the decoded record lands in binaryKmers (kmerBits longs, Java convention), coverages and edges
     *  (numColors each) at slot i. */
    static native void findRecords(long handle, byte[] kmers, int nq, long[] outIndex, long[] binaryKmers, int[] coverages, byte[] edges);

    // K3+K4 ------------------------------------------------------------------------------------ cc_find_ascii / cc_find_windows / cc_contains_windows
    /** nq k-byte ASCII k-mers, row-major -> record index per query, -1 = null. */
    static native void findAscii(long handle, byte[] kmers, int nq, long[] outIndex);
    /** Every k-window of seq -> record index (Call.loadChildWalk). */
    static native void findWindows(long handle, byte[] seq, long[] outIndex);
    /** Every k-window of seq -> present in this graph (ROI membership). */
    static native void containsWindows(long handle, byte[] seq, boolean[] outPresent);
    /** cc_pack_canonical: canonical packed words (Java long[] convention) and flags for every window of seq. */
    static native void packCanonical(int device, byte[] seq, int kmerSize, long[] outBinaryKmers, byte[] outFlags);

    // next rows -------------------------------------------------------------------------------- cc_join / cc_sort / cc_write_graph
    /** Each returns the handle of a NEW device-resident graph (dispose it). */
    static native long join(long[] handles);
    static native long sort(long handle);
    /** cc_remove: {handle of the new graph, merged records removed}. */
    static native long[] remove(long primaryHandle, long[] secondaryHandles);
    static native void writeGraph(long handle, String outPath);

    // pre-filters / recovery ------------------------------------------------------------------- cc_find_low_coverage / cc_find_shared / ...
    static native long findLowCoverage(long roiHandle, int minCoverage);
    static native long findShared(long graphHandle, long roiHandle, int child, int[] parents, int[] ignore);
    static native long recoverExcludedKmers(long graphHandle, long dirtyHandle, int child);
    /** {coverage0, count0, coverage1, count1, ...} in ascending coverage (cc_cov_stats). */
    static native int[] covStats(long handle, int child, int[] parents);

    // one graph over several GPUs ---------------------------------------------------------------- cc_open_sharded / cc_*_sharded
    /** One graph over several devices (ArgumentHandler.java:271-274 constructs ONE graph per file).  placement: 0 = k-mer-range shards,
     *  1 = a replica per device (lookups need no exchange), 2 = replicas when they fit, ranges otherwise (cc_open_sharded_placed). */
    static native long openSharded(String path, int[] devices, int placement);
    static native void disposeSharded(long shardedHandle);
    /** {numShards, numRecords, kmerSize, numColors} */
    static native long[] shardedInfo(long shardedHandle);
    /** {borrowed cc_graph handle of the shard (header, colours, decodeRecords), device, first record index} */
    static native long[] shardedShard(long shardedHandle, int rank);
    static native void findAsciiSharded(long shardedHandle, byte[] kmers, int nq, long[] outIndex);
    static native void findWindowsSharded(long shardedHandle, byte[] seq, long[] outIndex);
    static native long findNovelSharded(long shardedHandle, int child, int[] parents, byte[] outRecords, long[] outIndex);
    static native long writeRoiFileSharded(long shardedHandle, int child, int[] parents, String outPath);
}
