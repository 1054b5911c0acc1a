// B200MasterPlan.java -- java.lang.foreign (Panama, JDK 22+) shim that puts libb200llama.so behind
// the reference's TornadoVMMasterPlan interface (tornadovm/TornadoVMMasterPlan.java:30-85).
//
// SHIPPED AS SOURCE: this image has no JDK, so the file is not compiled or exercised here; the
// identical C ABI (include/b200llama.h) is exercised from Python ctypes in tests/.  Drop it into
// src/main/java/org/beehive/gpullama3/tornadovm/ of the reference (see INTEGRATION.md).
package org.beehive.gpullama3.tornadovm;

import org.beehive.gpullama3.inference.state.State;
import org.beehive.gpullama3.model.Configuration;
import org.beehive.gpullama3.model.Model;
import org.beehive.gpullama3.tensor.GGMLTensorEntry;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.util.Map;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/** One native plan = one TornadoVMMasterPlan: create, forward*, free. Single-owner, like the reference. */
public final class B200MasterPlan implements AutoCloseable {

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("b200.lib", "libb200llama.so"), Arena.global());

    // struct b200_config: 9 x int32, 2 x float, 3 x int32
    private static final StructLayout CONFIG = MemoryLayout.structLayout(
            JAVA_INT.withName("arch"), JAVA_INT.withName("dim"), JAVA_INT.withName("hidden_dim"), JAVA_INT.withName("n_layers"),
            JAVA_INT.withName("n_heads"), JAVA_INT.withName("n_kv_heads"), JAVA_INT.withName("head_size"), JAVA_INT.withName("vocab_size"),
            JAVA_INT.withName("context_length"), JAVA_FLOAT.withName("rms_norm_eps"), JAVA_FLOAT.withName("rope_theta"),
            JAVA_INT.withName("fp16_lanes"), JAVA_INT.withName("tp_rank"), JAVA_INT.withName("tp_size"));
    // struct b200_tensor: char* name, void* data, int32 type, int32 n_dims, int64 dims[4]
    private static final StructLayout TENSOR = MemoryLayout.structLayout(
            ADDRESS.withName("name"), ADDRESS.withName("data"), JAVA_INT.withName("ggml_type"), JAVA_INT.withName("n_dims"),
            MemoryLayout.sequenceLayout(4, JAVA_LONG).withName("dims"));

    private static MethodHandle fn(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), fd);
    }

    private static final MethodHandle CREATE = fn("b200_plan_create",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle DECODE = fn("b200_forward_decode", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle PREFILL = fn("b200_forward_prefill", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
    private static final MethodHandle BATCH_PREFILL = fn("b200_forward_batch_prefill", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
    private static final MethodHandle SET_PREFILL_MODE = fn("b200_set_prefill_mode", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle SET_DECODE_MODE = fn("b200_set_decode_mode", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle DECODE_SAMPLE = fn("b200_forward_decode_sample",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_FLOAT, JAVA_FLOAT, JAVA_FLOAT, ADDRESS, ADDRESS));
    private static final MethodHandle DECODE_SEQUENCE = fn("b200_decode_sequence",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle TP_HANDLE = fn("b200_tp_handle", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle TP_ATTACH = fn("b200_tp_attach", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle KV_RESET = fn("b200_kv_reset", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle FREE = fn("b200_plan_free", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle LAST_ERROR = fn("b200_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));

    private final Arena arena = Arena.ofConfined();
    private final MemorySegment plan;
    private final MemorySegment logits;   // vocab floats, reused every token (state.wrapLogits in the reference)
    private final MemorySegment argmax;

    /** TornadoVMMasterPlan.initializeTornadoVMPlan(state, model): tensors are the plain mmap slices of GGUF.loadTensorsStandard.
     *  archId: 0 = Llama / Mistral, 1 = Qwen3, 2 = Phi-3 (pass blk.N.attn_qkv.weight and blk.N.ffn_up.weight fused, as in the file). */
    public B200MasterPlan(State state, Model model, Map<String, GGMLTensorEntry> tensors, int archId, int headSize) throws Throwable {
        this(state, model, tensors, archId, headSize, 0, 1);
    }

    /** One plan per GPU for tensor-parallel decode: every rank passes the same tensors, the library uploads its row slices. */
    public B200MasterPlan(State state, Model model, Map<String, GGMLTensorEntry> tensors, int archId, int headSize, int tpRank, int tpSize) throws Throwable {
        Configuration c = model.configuration();
        MemorySegment cfg = arena.allocate(CONFIG);
        int[] ints = {archId, c.dim(), c.hiddenDim(), c.numberOfLayers(), c.numberOfHeads(), c.numberOfKeyValueHeads(), headSize,
                c.vocabularySize(), c.contextLength()};
        for (int i = 0; i < ints.length; i++) cfg.setAtIndex(JAVA_INT, i, ints[i]);
        cfg.set(JAVA_FLOAT, 36, c.rmsNormEps());
        cfg.set(JAVA_FLOAT, 40, c.ropeTheta());
        cfg.set(JAVA_INT, 44, Integer.getInteger("llama.VectorBitSize", 512) / 32); // FloatTensor.java:21
        cfg.set(JAVA_INT, 48, tpRank);
        cfg.set(JAVA_INT, 52, tpSize);

        MemorySegment arr = arena.allocate(TENSOR, tensors.size());
        int i = 0;
        for (var e : tensors.entrySet()) {
            MemorySegment t = arr.asSlice((long) i * TENSOR.byteSize(), TENSOR.byteSize());
            t.set(ADDRESS, 0, arena.allocateFrom(e.getKey()));
            t.set(ADDRESS, 8, e.getValue().memorySegment());       // MemorySegment.address() of the mapping
            t.set(JAVA_INT, 16, e.getValue().ggmlType().ordinal()); // GGMLType ordinal == ggml type id (F32 0, F16 1, Q8_0 8, Q4_K 12, Q5_K 13, Q6_K 14: GGMLType.java:5-20); K-quants are re-quantised on the device
            int[] shape = e.getValue().shape();
            t.set(JAVA_INT, 20, shape.length);
            for (int d = 0; d < shape.length; d++) t.set(JAVA_LONG, 24 + 8L * d, shape[d]);
            i++;
        }
        MemorySegment out = arena.allocate(ADDRESS);
        MemorySegment err = arena.allocate(512);
        int batch = TornadoVMMasterPlan.WITH_PREFILL_DECODE ? TornadoVMMasterPlan.PREFILL_BATCH_SIZE : 0;
        int rc = (int) CREATE.invokeExact(cfg, arr, tensors.size(), batch, Integer.getInteger("b200.device", tpRank), out, err, 512L);
        check(rc, err.getString(0));
        plan = out.get(ADDRESS, 0);
        logits = arena.allocate(JAVA_FLOAT, c.vocabularySize());
        argmax = arena.allocate(JAVA_INT);
    }

    private static void check(int rc, String msg) {
        if (rc == 0) return;
        if (rc == -2) throw new UnsupportedOperationException(msg);            // ForwardPlanFactory.java:84-87
        if (rc == -3) throw new OutOfMemoryError("B200 device memory: " + msg); // README.md:262-265
        throw new IllegalStateException("b200llama error " + rc + ": " + msg);
    }

    private String lastError() throws Throwable {
        return ((MemorySegment) LAST_ERROR.invokeExact(plan)).reinterpret(512).getString(0);
    }

    /** FloatArray tornadoVMForwardDecode(int position) + the embedding gather of InferenceCore.forwardTornadoVM (InferenceCore.java:956-980). */
    public MemorySegment forwardDecode(int token, int position) throws Throwable {
        int rc = (int) DECODE.invokeExact(plan, token, position, logits, argmax);
        if (rc != 0) check(rc, lastError());
        return logits;
    }

    /** Greedy path: only the argmax (4 bytes) crosses PCIe. */
    public int forwardDecodeArgmax(int token, int position) throws Throwable {
        int rc = (int) DECODE.invokeExact(plan, token, position, MemorySegment.NULL, argmax);
        if (rc != 0) check(rc, lastError());
        return argmax.get(JAVA_INT, 0);
    }

    /** void tornadoVMForwardPrefill(int position) (TornadoVMMasterPlanPrefillDecode.java:116). */
    public void forwardPrefill(int token, int position) throws Throwable {
        int rc = (int) PREFILL.invokeExact(plan, token, position);
        if (rc != 0) check(rc, lastError());
    }

    /** void tornadoVMForwardBatchPrefill() (TornadoVMMasterPlanBatchPrefillDecode.java:107-123). */
    public void forwardBatchPrefill(int[] tokens, int startPos) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            int rc = (int) BATCH_PREFILL.invokeExact(plan, a.allocateFrom(JAVA_INT, tokens), tokens.length, startPos);
            if (rc != 0) check(rc, lastError());
        }
    }

    /** Forward + Sampler.sampleToken on the device (Sampler.java:74-122): the caller draws uniform01 = rng.nextFloat(1f) from the
     *  sampler's own RandomGenerator (Sampler.java:84) and gets the token id back; the logits row never crosses PCIe. */
    public int forwardDecodeSample(int token, int position, float temperature, float topp, float uniform01) throws Throwable {
        int rc = (int) DECODE_SAMPLE.invokeExact(plan, token, position, temperature, topp, uniform01, argmax, MemorySegment.NULL);
        if (rc != 0) check(rc, lastError());
        return argmax.get(JAVA_INT, 0);
    }

    /** 0 = CUDA graph of ~7 kernels per layer, 1 = one persistent kernel per token (default when the plan supports it); bit-identical. */
    public void setDecodeMode(int mode) throws Throwable {
        int rc = (int) SET_DECODE_MODE.invokeExact(plan, mode);
        if (rc != 0) check(rc, lastError());
    }

    /** TensorCoreSupport.java's switch: 0 = exact token-by-token prefill (bit-identical KV cache), 1 = TMA + tcgen05 GEMMs. */
    public void setPrefillMode(int mode) throws Throwable {
        int rc = (int) SET_PREFILL_MODE.invokeExact(plan, mode);
        if (rc != 0) check(rc, lastError());
    }

    /** The greedy loop of InferenceEngine.generateTokensGPULlama with sampler and token feedback on the device (feedback = true),
     *  or LlamaBench's teacher-forced loop (feedback = false): n steps from startPos, returns the argmax of every step. */
    public int[] decodeSequence(int[] tokens, int n, int startPos, boolean feedback) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment ids = a.allocate(JAVA_INT, n);
            int rc = (int) DECODE_SEQUENCE.invokeExact(plan, a.allocateFrom(JAVA_INT, tokens), n, startPos, feedback ? 1 : 0, ids, MemorySegment.NULL);
            if (rc != 0) check(rc, lastError());
            return ids.toArray(JAVA_INT);
        }
    }

    /** 64-byte CUDA-IPC handle of this rank's communication buffer; exchange with the other ranks, then attach(all handles in rank order). */
    public byte[] tpHandle() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment h = a.allocate(64);
            int rc = (int) TP_HANDLE.invokeExact(plan, h);
            if (rc != 0) check(rc, lastError());
            return h.toArray(java.lang.foreign.ValueLayout.JAVA_BYTE);
        }
    }

    public void tpAttach(byte[][] handles) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment all = a.allocate(64L * handles.length);
            for (int r = 0; r < handles.length; r++) MemorySegment.copy(handles[r], 0, all, java.lang.foreign.ValueLayout.JAVA_BYTE, 64L * r, 64);
            int rc = (int) TP_ATTACH.invokeExact(plan, all, handles.length);
            if (rc != 0) check(rc, lastError());
        }
    }

    public void kvReset() throws Throwable {
        int rc = (int) KV_RESET.invokeExact(plan);
        if (rc != 0) check(rc, lastError());
    }

    /** void freeTornadoExecutionPlan() */
    @Override
    public void close() {
        try {
            FREE.invokeExact(plan);
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        } finally {
            arena.close();
        }
    }
}
