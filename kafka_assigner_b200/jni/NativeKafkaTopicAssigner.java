// NativeKafkaTopicAssigner.java — reference-side binding of libkassign.so (see INTEGRATION.md).
// NOT compiled in this repository: the build image has no JDK. Drop next to
// src/main/java/siftscience/kafka/tools/KafkaTopicAssigner.java of the reference and change
// KafkaAssignmentGenerator.java:172 to `new NativeKafkaTopicAssigner()`.
package siftscience.kafka.tools;

import java.util.*;

/** KafkaTopicAssigner whose generateAssignment runs on the GPU through libkassign.so. */
public class NativeKafkaTopicAssigner extends KafkaTopicAssigner implements AutoCloseable {
    static { System.loadLibrary("kassign_jni"); }

    private final long ctx = NativeAssigner.create(0);          // one Context per instance, like KTA:19-23
    private int[] brokerIds;                                    // sorted; re-uploaded only when the set changes
    private Map<Integer, String> lastRacks;

    @Override
    public Map<Integer, List<Integer>> generateAssignment(String topic, Map<Integer, List<Integer>> cur,
            Set<Integer> brokers, Map<Integer, String> racks, int desiredRf) {
        if (ctx == 0) throw new IllegalStateException("kassign: no CUDA device (there is no CPU fallback)");
        int[] ids = brokers.stream().mapToInt(Integer::intValue).sorted().toArray();
        if (!Arrays.equals(ids, brokerIds) || !racks.equals(lastRacks)) {
            String[] names = new String[ids.length];
            for (int i = 0; i < ids.length; i++) names[i] = racks.get(ids[i]);   // null = no rack (KAS:82-86)
            NativeAssigner.setBrokers(ctx, ids, names);
            brokerIds = ids; lastRacks = new HashMap<>(racks);
        }
        int[] parts = cur.keySet().stream().mapToInt(Integer::intValue).sorted().toArray();   // TreeMap order KAS:107-110
        long[] repOff = new long[parts.length + 1];
        int maxLen = 0;
        for (int i = 0; i < parts.length; i++) { int n = cur.get(parts[i]).size(); repOff[i + 1] = repOff[i] + n; maxLen = Math.max(maxLen, n); }
        int[] flat = new int[(int) repOff[parts.length]];
        for (int i = 0, k = 0; i < parts.length; i++) for (int b : cur.get(parts[i])) flat[k++] = b;
        int stride = Math.max(1, Math.max(maxLen, Math.max(desiredRf, 0)));
        int[] outLen = new int[parts.length], out = new int[parts.length * stride];
        // batch of ONE topic; throws the reference's own exceptions with identical messages
        NativeAssigner.solve(ctx, new String[]{topic}, new int[]{topic.hashCode()}, new long[]{0, parts.length}, parts,
                             repOff, flat, desiredRf, stride, outLen, out);
        Map<Integer, List<Integer>> res = new TreeMap<>();
        for (int i = 0; i < parts.length; i++) {
            List<Integer> l = new ArrayList<>(outLen[i]);
            for (int r = 0; r < outLen[i]; r++) l.add(out[i * stride + r]);
            res.put(parts[i], l);
        }
        return res;
    }
    @Override public void close() { NativeAssigner.destroy(ctx); }
}

final class NativeAssigner {
    static native long create(int device);
    static native void destroy(long ctx);
    static native void setBrokers(long ctx, int[] sortedIds, String[] rackNamesOrNull);
    static native void solve(long ctx, String[] topicNames, int[] topicHash, long[] partOff, int[] partId, long[] repOff,
                             int[] curBroker, int desiredRf, int outStride, int[] outLen, int[] outBroker);
}
