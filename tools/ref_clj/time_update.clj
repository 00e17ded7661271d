;; Times one mcl* iteration of the REAL reference (jvia/amclj) — to be loaded from `lein repl`
;; inside a checkout of the reference.  UNVERIFIED: no JVM exists where this file was written.
;; It only calls the reference's public functions; nothing of the reference is copied here.
(ns amclj-b200.ref-timing
  (:require [amclj.pf :as pf]
            [geometry-msgs]
            [nav-msgs]
            [sensor-msgs]
            [std-msgs]
            [incanter.core :refer [matrix]]))

(defn- read-cells [path width]
  (let [f (java.io.File. path)
        bytes (byte-array (.length f))]
    (with-open [in (java.io.DataInputStream. (java.io.FileInputStream. f))]
      (.readFully in bytes))
    (matrix (vec bytes) width)))                      ; nav_msgs.clj:38 builds :data the same way

(defn- timed [label thunk]
  (let [t0 (System/nanoTime)
        v (thunk)
        ms (/ (- (System/nanoTime) t0) 1e6)]
    (println (format "%-22s %10.1f ms" label ms))
    [v ms]))

(defn run [cells-path n-particles n-beams]
  (let [width 662 height 639
        grid (nav-msgs/occupancy-grid
              :info (nav-msgs/map-meta-data :resolution (float 0.05) :width width :height height)
              :data (read-cells cells-path width))
        pose (geometry-msgs/pose :position (geometry-msgs/point :x 16.3 :y 15.8)
                                 :orientation (geometry-msgs/quaternion :z 0.85 :w 0.519))
        cloud (pf/initialize-pf grid n-particles :pose pose)
        scan (sensor-msgs/laser-scan :angleMin (float (- (/ Math/PI 2)))
                                     :angleIncrement (float (/ Math/PI (dec n-beams)))
                                     :rangeMax (float 5.6)
                                     :ranges (vec (repeat n-beams (float 2.0))))
        tf (fn [x y] (geometry-msgs/transform :translation (geometry-msgs/vector3 :x x :y y)
                                              :rotation (geometry-msgs/quaternion :w 1.0)))
        control [(tf 0.2 0.0) (tf 0.0 0.0)]
        [moved t-motion] (timed "apply-motion-model" #(update-in (pf/apply-motion-model control cloud) [:poses] doall))
        [weighted t-sensor] (timed "apply-sensor-model" #(update-in (pf/apply-sensor-model scan grid moved) [:poses] doall))
        [_ t-select] (timed "roulette-selection" #(update-in (pf/roulette-selection weighted) [:poses] doall))]
    (println (format "beam evaluations / s (30 beams per particle): %.3e"
                     (/ (* n-particles 30.0) (/ t-sensor 1e3))))
    {:motion-ms t-motion :sensor-ms t-sensor :selection-ms t-select}))
