(ns amclj.gpu
  "B200 back end for amclj's filter hot path: JNA binding of libamclj_b200.so (include/amclj_b200.h).

  Replaces the bodies of pf/apply-motion-model (pf.clj:139-145), pf/apply-sensor-model (:148-154) and
  pf/roulette-selection (:156-167) — see shim/pf.clj.patch — with calls that keep the particle set on
  the GPU.  The record shapes at the boundary are the reference's own: geometry_msgs/Pose
  (geometry_msgs.clj:54-72) as rows of 7 doubles (position x y z, orientation x y z w),
  nav_msgs/OccupancyGrid (nav_msgs.clj:25-38) and sensor_msgs/LaserScan (sensor_msgs.clj:8-20).
  quat/to-heading and quat/from-heading run INSIDE the library (amclj_set_poses_quat /
  amclj_get_poses_quat), so this namespace does no arithmetic of its own.

  UNVERIFIED: written against the C header; the build image has no JVM, so this file has never been
  loaded.  The same entry points are exercised from C (tests/c/*.c) and Python ctypes (amclj_b200/_lib.py).

  project.clj: add [net.java.dev.jna/jna \"4.1.0\"] to :dependencies; run with
  JNA_LIBRARY_PATH=<repo>/amclj_b200."
  (:require [geometry-msgs]
            [incanter.core :as ic])
  (:import [com.sun.jna Function Memory Pointer NativeLibrary]
           [com.sun.jna.ptr PointerByReference]))

(def ^:private lib (delay (NativeLibrary/getInstance "amclj_b200")))
(defn- f ^Function [name] (.getFunction ^NativeLibrary @lib name))

(def preset-ref 0) ; the reference as written, bug for bug (30 beams, -1.0 m rays, sum p^3, roulette)
(def preset-ns 1)  ; north-star semantics (heading-correct rays, all beams, log-weights, systematic)
(def resample-roulette 0)
(def resample-systematic 1)
(def resample-tournament 2)

(defn call!
  "Invoke an int32-status entry point on a ctx; a negative status throws with amclj_last_error."
  [ctx name & args]
  (let [rc (.invokeInt (f name) (to-array (cons ctx args)))]
    (when (neg? rc)
      (throw (ex-info (.invokeString (f "amclj_last_error") (to-array [ctx]) false)
                      {:code rc :call name})))
    rc))

(defn create-ctx
  "One ctx per GPU.  Fails with AMCLJ_ERR_NODEVICE (-5) when no CUDA device is usable: there is no
  CPU fallback."
  [device preset]
  (let [out (PointerByReference.)
        rc (.invokeInt (f "amclj_create") (to-array [(int device) out]))]
    (when (neg? rc) (throw (ex-info "amclj_create failed" {:code rc})))
    (let [ctx (.getValue out)
          params (Memory. 112)]                       ; sizeof(amclj_params), checked by the C tests
      (.invokeInt (f "amclj_params_default") (to-array [params (int preset)]))
      (call! ctx "amclj_set_params" params)
      ctx)))

(defn destroy-ctx [ctx] (.invokeInt (f "amclj_destroy") (to-array [ctx])))

(defn- float-mem ^Memory [xs]
  (let [a (float-array xs) m (Memory. (max 4 (* 4 (alength a))))]
    (.write m 0 a 0 (alength a)) m))

(defn- double-mem ^Memory [xs]
  (let [a (double-array xs) m (Memory. (max 8 (* 8 (alength a))))]
    (.write m 0 a 0 (alength a)) m))

(defn set-map!
  "nav_msgs/OccupancyGrid as amclj holds it (nav_msgs.clj:27-38): :data is a height x width matrix,
  -1 unknown / 0 free / 100 occupied; :info :resolution is a ROS float32."
  [ctx grid]
  (let [{:keys [width height resolution]} (:info grid)
        cells (byte-array (map #(byte (int %)) (flatten (ic/to-list (:data grid)))))
        m (doto (Memory. (alength cells)) (.write 0 cells 0 (alength cells)))]
    (call! ctx "amclj_set_map" m (int width) (int height) (float resolution))))

(defn- pose-rows
  "PoseArray :poses -> [n][7] doubles in the record's own field order."
  [poses]
  (mapcat (fn [p] [(get-in p [:position :x]) (get-in p [:position :y]) (get-in p [:position :z] 0.0)
                   (get-in p [:orientation :x]) (get-in p [:orientation :y])
                   (get-in p [:orientation :z]) (get-in p [:orientation :w])])
          poses))

(defn upload-poses!
  "PoseArray :poses -> device.  theta = quat/to-heading of every orientation, computed by the library
  exactly as quaternions.clj:82-97 (pole guards included)."
  [ctx poses]
  (call! ctx "amclj_set_poses_quat" (double-mem (pose-rows poses)) Pointer/NULL (long (count poses))))

(defn download-poses
  "Device -> :poses with :weight.  orientation = quat/from-heading(theta) (quaternions.clj:57-62)."
  [ctx]
  (let [n (.invokeLong (f "amclj_num_particles") (to-array [ctx]))
        rows (Memory. (* 56 n))
        w (Memory. (* 4 n))]
    (call! ctx "amclj_get_poses_quat" rows w)
    (vec (for [i (range n)]
           (let [d #(.getDouble rows (+ (* 56 i) (* 8 %)))]
             (assoc (geometry-msgs/pose
                     :position (geometry-msgs/point :x (d 0) :y (d 1) :z (d 2))
                     :orientation (geometry-msgs/quaternion :x (d 3) :y (d 4) :z (d 5) :w (d 6)))
               :weight (.getFloat w (* 4 i))))))))

(defn- yaw-pose
  "A tf transform -> (x, y, yaw) doubles; yaw through the library's to-heading by way of a one-row upload
  would be wasteful, so the reference's own function is used here (motion.clj:76-79 does the same)."
  [tf to-heading]
  (double-mem [(-> tf :translation :x) (-> tf :translation :y) (to-heading (:rotation tf))]))

(defn motion-update!
  "pf.clj:139-145.  control = [curr-tf prev-tf]; NULL normals = in-kernel Philox keyed by seed."
  [ctx [curr prev] to-heading seed]
  (call! ctx "amclj_motion_update" (yaw-pose curr to-heading) (yaw-pose prev to-heading)
         Pointer/NULL (long seed)))

(defn sensor-update!
  "pf.clj:148-154.  LaserScan fields as sensor_msgs.clj:8-12 names them."
  [ctx scan]
  (let [r (:ranges scan)]
    (call! ctx "amclj_sensor_update" (float-mem r) (int (count r))
           (float (:angleMin scan)) (float (:angleIncrement scan)) (float (:rangeMax scan)))))

(defn resample!
  "pf.clj:156-167 (mode 0), low-variance (1), pf.clj:182-195 (2)."
  [ctx mode u0 seed]
  (call! ctx "amclj_resample" (int mode) (double u0) Pointer/NULL Pointer/NULL (long 0) (long seed)
         Pointer/NULL))

(defn filter-update!
  "One mcl* iteration's hot path (pf.clj:229-233) as ONE device-resident call."
  [ctx [curr prev] to-heading scan u0 seed]
  (let [r (:ranges scan)]
    (call! ctx "amclj_filter_update" (yaw-pose curr to-heading) (yaw-pose prev to-heading)
           Pointer/NULL (long seed) (float-mem r) (int (count r))
           (float (:angleMin scan)) (float (:angleIncrement scan)) (float (:rangeMax scan)) (double u0))))

(defn filter-update-async!
  "The same iteration without waiting for it: scan staging and update in one call (the motion kernel is enqueued
  first and the message's beam table built while it runs once the laser's geometry is staged).  Systematic
  resampling only; (call! ctx \"amclj_synchronize\") or any download waits for it."
  [ctx [curr prev] to-heading scan u0 seed]
  (let [r (:ranges scan)]
    (call! ctx "amclj_filter_update_scan_async" (yaw-pose curr to-heading) (yaw-pose prev to-heading)
           (long seed) (double u0) (float-mem r) (int (count r))
           (float (:angleMin scan)) (float (:angleIncrement scan)) (float (:rangeMax scan)))))

(defn host-buffers
  "Four float buffers (x, y, theta, w) of n particles for filter-update-host!, page-locked once with
  amclj_host_register: a JNA Memory is plain malloc'ed memory, and copies from pageable memory go through the
  driver's staging buffers at about a quarter of the bus's speed.  Keep and reuse them; release-host-buffers!
  before they are garbage-collected."
  [ctx n]
  (let [bufs (vec (repeatedly 4 #(Memory. (* 4 (long n)))))]
    (doseq [^Memory b bufs] (call! ctx "amclj_host_register" b (long (.size b))))
    bufs))

(defn release-host-buffers! [ctx bufs]
  (doseq [^Memory b bufs] (call! ctx "amclj_host_unregister" b)))

(defn filter-update-host!
  "The same iteration for a host-resident cloud (the reference keeps its particles in the JVM): x, y, theta in
  the buffers of host-buffers are uploaded, updated and overwritten by the resampled cloud and its weights."
  [ctx [x y theta w] n [curr prev] to-heading scan u0 seed]
  (let [r (:ranges scan)]
    (call! ctx "amclj_filter_update_host" x y theta w (long n)
           (yaw-pose curr to-heading) (yaw-pose prev to-heading) Pointer/NULL (long seed)
           (float-mem r) (int (count r))
           (float (:angleMin scan)) (float (:angleIncrement scan)) (float (:rangeMax scan)) (double u0))))

(defn pose-estimate
  "pf.clj:67-84: unweighted mean of x, y and of the quaternion's z, w, without a download."
  [ctx]
  (let [out (Memory. 32)]
    (call! ctx "amclj_pose_estimate" out)
    (geometry-msgs/pose-stamped
     :pose (geometry-msgs/pose
            :position (geometry-msgs/point :x (.getDouble out 0) :y (.getDouble out 8))
            :orientation (geometry-msgs/quaternion :x 0.0 :y 0.0 :z (.getDouble out 16) :w (.getDouble out 24))))))

;; ---- several GPUs from the one caller thread the reference has (pf.clj:297): amclj_mctx -------------
(defn mcall! [m name & args]
  (let [rc (.invokeInt (f name) (to-array (cons m args)))]
    (when (neg? rc)
      (throw (ex-info (.invokeString (f "amclj_mlast_error") (to-array [m]) false) {:code rc :call name})))
    rc))

(defn create-multi
  "devices: CUDA ordinals, one shard of the particle set per entry."
  [devices preset]
  (let [out (PointerByReference.)
        d (int-array devices)
        dm (doto (Memory. (* 4 (alength d))) (.write 0 d 0 (alength d)))
        rc (.invokeInt (f "amclj_create_multi") (to-array [dm (int (alength d)) out]))]
    (when (neg? rc) (throw (ex-info "amclj_create_multi failed" {:code rc})))
    (let [m (.getValue out) params (Memory. 112)]
      (.invokeInt (f "amclj_params_default") (to-array [params (int preset)]))
      (mcall! m "amclj_mset_params" params)
      m)))

(defn mfilter-update!
  "The same iteration over all GPUs: the kernels exchange {max, total} through peer-memory mailboxes
  and pull their resampled slots over NVLink; the host only enqueues."
  [m [curr prev] to-heading scan u0 seed]
  (let [r (:ranges scan)]
    (mcall! m "amclj_mstage_scan" (float-mem r) (int (count r))
            (float (:angleMin scan)) (float (:angleIncrement scan)) (float (:rangeMax scan)))
    (mcall! m "amclj_mfilter_update_async" (yaw-pose curr to-heading) (yaw-pose prev to-heading)
            (long seed) (double u0))
    (mcall! m "amclj_msynchronize")))
