package mergingBodies3D;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;

/**
 * Panama FFM binding of libam3d.so (include/am3d.h).  One instance = one am3d_ctx = one GPU.
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no JDK there); needs JDK >= 22 (java.lang.foreign is final).
 * It lives in package mergingBodies3D so that it can read the package-private fields of the reference classes.
 * Every handle below binds one declaration of include/am3d.h; the Python ctypes binding used by the tests
 * (adaptivemerging_b200/_capi.py) makes the same calls.
 */
final class AM3DNative implements AutoCloseable {
    private static final Linker L = Linker.nativeLinker();
    private static final SymbolLookup LIB =
            SymbolLookup.libraryLookup(System.getProperty("am3d.library", "libam3d.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor d) {
        return L.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
    }

    private static final MethodHandle CREATE = h("am3d_create", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle DESTROY = h("am3d_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle LAST_ERROR = h("am3d_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle DEFAULT_PARAMS = h("am3d_default_params", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle UPLOAD_SCENE = h("am3d_upload_scene", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle SET_PARAMS = h("am3d_set_params", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle RESET = h("am3d_reset", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle STEP = h("am3d_step", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_DOUBLE, JAVA_INT));
    private static final MethodHandle NUM_BODIES = h("am3d_num_bodies", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle DOWNLOAD_BODIES = h("am3d_download_bodies",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle DOWNLOAD_BODIES_ASYNC = h("am3d_download_bodies_async",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle WAIT_DOWNLOAD = h("am3d_wait_download", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle NUM_CONTACTS = h("am3d_num_contacts", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle DOWNLOAD_CONTACTS = h("am3d_download_contacts",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle GET_TIMINGS = h("am3d_get_timings", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle SET_BODY_VELOCITY = h("am3d_set_body_velocity",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle ADD_BODY_VELOCITY = h("am3d_add_body_velocity",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle ADD_VELOCITIES = h("am3d_add_velocities", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    // callers' hooks (SURVEY.md 8f.2): Factory / add / remove, mouse tools, Animation, list order, async stepping
    private static final MethodHandle ACTIVATE_BODY = h("am3d_activate_body", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle REMOVE_BODY = h("am3d_remove_body", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle SET_MOUSE_SPRING = h("am3d_set_mouse_spring",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_INT));
    private static final MethodHandle APPLY_IMPULSE = h("am3d_apply_impulse", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_DOUBLE));
    private static final MethodHandle SET_BODY_SLEEPING = h("am3d_set_body_sleeping", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
    private static final MethodHandle SET_BODY_MAGNET = h("am3d_set_body_magnet", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
    private static final MethodHandle LIST_ORDER = h("am3d_download_list_order", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle STEP_ASYNC = h("am3d_step_async", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_DOUBLE, JAVA_INT));
    private static final MethodHandle SYNC = h("am3d_sync", FunctionDescriptor.of(JAVA_INT, ADDRESS));

    /** sizeof(am3d_params), sizeof(am3d_timings), sizeof(am3d_contact): see the struct definitions in am3d.h */
    static final long SIZEOF_PARAMS = 264, SIZEOF_TIMINGS = 128, SIZEOF_CONTACT = 264;

    final Arena arena = Arena.ofShared();
    private final MemorySegment ctx;

    AM3DNative(int device) {
        MemorySegment out = arena.allocate(ADDRESS);
        // AM3D_ENOGPU (-3) when there is no CUDA device: the library has no CPU fallback
        int rc;
        try { rc = (int) CREATE.invokeExact(device, out); } catch (Throwable t) { throw new RuntimeException(t); }
        if (rc != 0) throw new IllegalStateException("am3d_create failed: " + rc);
        ctx = out.get(ADDRESS, 0);
    }

    private void check(int rc) {
        if (rc == 0) return;
        String msg;
        try {
            MemorySegment s = (MemorySegment) LAST_ERROR.invokeExact(ctx);
            msg = s.reinterpret(4096).getString(0);
        } catch (Throwable t) { msg = "?"; }
        throw new IllegalStateException("am3d error " + rc + ": " + msg);
    }

    MemorySegment defaultParams() {
        MemorySegment p = arena.allocate(SIZEOF_PARAMS, 8);
        try { DEFAULT_PARAMS.invokeExact(p); } catch (Throwable t) { throw new RuntimeException(t); }
        return p;
    }
    void uploadScene(MemorySegment am3dScene) { try { check((int) UPLOAD_SCENE.invokeExact(ctx, am3dScene)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void setParams(MemorySegment am3dParams) { try { check((int) SET_PARAMS.invokeExact(ctx, am3dParams)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void reset() { try { check((int) RESET.invokeExact(ctx)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void step(double dt, int n) { try { check((int) STEP.invokeExact(ctx, dt, n)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    int numBodies() { try { return (int) NUM_BODIES.invokeExact(ctx); } catch (Throwable t) { throw new RuntimeException(t); } }
    int numContacts(boolean includeInternal) { try { return (int) NUM_CONTACTS.invokeExact(ctx, includeInternal ? 1 : 0); } catch (Throwable t) { throw new RuntimeException(t); } }

    /** x[3n] R[9n] v[3n] omega[3n] (double), sleeping[n] collection[n] (int): caller-owned off-heap buffers */
    void downloadBodies(MemorySegment x, MemorySegment R, MemorySegment v, MemorySegment w, MemorySegment sleeping, MemorySegment collection) {
        try { check((int) DOWNLOAD_BODIES.invokeExact(ctx, x, R, v, w, sleeping, collection)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }
    /** same read, copied out on a second stream while the next step runs; valid after waitDownload() (draw one frame behind) */
    void downloadBodiesAsync(MemorySegment x, MemorySegment R, MemorySegment v, MemorySegment w, MemorySegment sleeping, MemorySegment collection) {
        try { check((int) DOWNLOAD_BODIES_ASYNC.invokeExact(ctx, x, R, v, w, sleeping, collection)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }
    void waitDownload() { try { check((int) WAIT_DOWNLOAD.invokeExact(ctx)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    /** out: capacity x am3d_contact; returns the number written */
    int downloadContacts(MemorySegment out, int capacity, boolean includeInternal) {
        MemorySegment n = arena.allocate(JAVA_INT);
        try { check((int) DOWNLOAD_CONTACTS.invokeExact(ctx, out, capacity, includeInternal ? 1 : 0, n)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
        return n.get(JAVA_INT, 0);
    }
    void timings(MemorySegment am3dTimings) { try { check((int) GET_TIMINGS.invokeExact(ctx, am3dTimings)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    /** MouseImpulse / scripted pushes: v or w may be MemorySegment.NULL (leave unchanged) */
    void setBodyVelocity(int body, MemorySegment v3, MemorySegment w3) { try { check((int) SET_BODY_VELOCITY.invokeExact(ctx, body, v3, w3)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void addBodyVelocity(int body, MemorySegment dv3, MemorySegment dw3) { try { check((int) ADD_BODY_VELOCITY.invokeExact(ctx, body, dv3, dw3)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void addVelocities(MemorySegment dv, MemorySegment dw) { try { check((int) ADD_VELOCITIES.invokeExact(ctx, dv, dw)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }

    /** Factory.generateBody (Factory.java:99-116): a dormant clone (body_flags bit 16 in the scene blob) enters system.bodies */
    void activateBody(int body, MemorySegment x3, MemorySegment R9, MemorySegment v3, MemorySegment w3) { try { check((int) ACTIVATE_BODY.invokeExact(ctx, body, x3, R9, v3, w3)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void removeBody(int body) { try { check((int) REMOVE_BODY.invokeExact(ctx, body)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    /** MouseSpringForce.setPicked + the stiffness / damping parameters; body = -1 releases */
    void setMouseSpring(int body, MemorySegment grabPointB, MemorySegment pointW, double k, double c, boolean atCOM) { try { check((int) SET_MOUSE_SPRING.invokeExact(ctx, body, grabPointB, pointW, k, c, atCOM ? 1 : 0)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    /** MouseImpulse.release(): applied inside the next step */
    void applyImpulse(int body, MemorySegment pickedPointB, MemorySegment endPointW, double scale) { try { check((int) APPLY_IMPULSE.invokeExact(ctx, body, pickedPointB, endPointW, scale)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    /** RigidBody.activateMagnet as LCPApp3D's key 7 toggles it (LCPApp3D.java:936-947) */
    void setBodyMagnet(int body, boolean active) { try { check((int) SET_BODY_MAGNET.invokeExact(ctx, body, active ? 1 : 0)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void setBodySleeping(int body, boolean sleeping) { try { check((int) SET_BODY_SLEEPING.invokeExact(ctx, body, sleeping ? 1 : 0)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    /** monotone position keys of every leaf's top-level entity: sort by them to rebuild system.bodies in the reference's order */
    void listOrder(MemorySegment int64PerBody) { try { check((int) LIST_ORDER.invokeExact(ctx, int64PerBody)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void stepAsync(double dt, int n) { try { check((int) STEP_ASYNC.invokeExact(ctx, dt, n)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }
    void sync() { try { check((int) SYNC.invokeExact(ctx)); } catch (RuntimeException e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); } }

    @Override public void close() {
        try { int rc = (int) DESTROY.invokeExact(ctx); } catch (Throwable ignored) { }
        arena.close();
    }
}
