"""CPU restatement of the reference's association / track lifecycle -- TEST INFRASTRUCTURE ONLY.

Pure-Python loops (small cases only), each step citing the reference line it follows
(MOT.cpp = src/multiple_object_tracking_lidar.cpp).  Floats are numpy scalars so that every operation rounds in the
precision the C++ expression has (float members, double locals).  The IHGP call goes through the C++ oracle
(orc_ihgp_step), one track at a time in this_objIDs order, exactly like callIHGP's loop.  Parity status: pinned -- tests/test_ref_pin.py
runs a 140-frame scenario through the reference's own cloudCallback (oracle/_ref) and through this class: ids, rings and
published rows agree frame by frame (golden copy in tests/golden/ref_vectors.npz).
"""
import numpy as np

from . import oracle

f32, f64 = np.float32, np.float64


class TrackerRef:
    def __init__(self, frequency, id_threshold, data_length, lpf_tau, hyp_x, hyp_y):
        self.frequency = f32(frequency)
        self.id_threshold = f32(id_threshold)
        self.L = int(data_length)
        self.lpf_tau = f32(lpf_tau)
        self.dt_gp = f32(1) / self.frequency                      # MOT.cpp:159 (policy: always 1/frequency)
        self.cx = oracle.ihgp_setup(float(self.dt_gp), *hyp_x)
        self.cy = oracle.ihgp_setup(float(self.dt_gp), *hyp_y)
        self.obj_ids, self.stack, self.m = [], [], []             # objIDs, stack_obj, GP state (MOT.h:106-115)
        self.next_obj_num = 0
        self.spin_counter = 0
        self.first = True

    def _register(self, c):                                        # registerNewObstacle, MOT.cpp:507-543
        self.obj_ids.append(self.next_obj_num)
        self.next_obj_num += 1
        self.stack.append([c.copy() for _ in range(self.L)])
        self.m.append(np.zeros(4))

    @staticmethod
    def _euc_dist(ax, ay, bx, by):                                 # euc_dist, MOT.cpp:1025-1028 (double math, float result)
        dx, dy = f64(ax) - f64(bx), f64(ay) - f64(by)
        return f32(np.sqrt(dx * dx + dy * dy + f64(0)))

    def _fill(self, i, c):                                         # fill_with_linear_interpolation, MOT.cpp:593-619
        last = self.stack[i][self.L - 1]
        dx_total = f64(f32(c[0]) - f32(last[0]))
        dy_total = f64(f32(c[1]) - f32(last[1]))
        dt_total = f64(f32(c[3]) - f32(last[3]))
        lost_num = int(np.floor(abs(dt_total / f64(self.dt_gp)) + 0.5) * np.sign(dt_total / f64(self.dt_gp))) - 1  # C round()
        for _ in range(lost_num):
            lc = self.stack[i][self.L - 1]
            center = np.array([f32(f64(lc[0]) + dx_total / f64(lost_num)), f32(f64(lc[1]) + dy_total / f64(lost_num)),
                               f32(f64(lc[2]) + f64(0) / f64(lost_num)), f32(lc[3]) + self.dt_gp], dtype=np.float32)
            self.stack[i].pop(0)
            self.stack[i].append(center)

    def step(self, centroids, now):
        """centroids: K x 4 float32 (x, y, z, intensity).  Returns None on the first frame / empty input, else
        (this_objIDs, pos_vel K x 8)."""
        centroids = np.asarray(centroids, dtype=np.float32).reshape(-1, 4)
        if len(centroids) == 0:                                    # MOT.cpp:146-150 / 170-174
            return None
        if self.first:                                             # MOT.cpp:126-161
            for c in centroids:
                self._register(c)
            self.first = False
            return None
        this_ids = []
        for c in centroids:                                        # MOT.cpp:177-219
            registered = False
            for index in range(len(self.obj_ids)):
                last = self.stack[index][self.L - 1]
                if self._euc_dist(c[0], c[1], last[0], last[1]) < self.id_threshold:
                    if f32(c[3]) - f32(last[3]) > f32(3) * self.dt_gp:
                        self._fill(index, c)
                    registered = True
                    break
            if registered:
                self.stack[index].pop(0)                           # updateObstacleQueue, MOT.cpp:586-591
                self.stack[index].append(c.copy())
                this_ids.append(self.obj_ids[index])
            else:
                this_ids.append(self.next_obj_num)
                self._register(c)
        pos_vel = np.zeros((len(this_ids), 8), dtype=np.float32)
        for e, n in enumerate(this_ids):                           # callIHGP, MOT.cpp:621-662
            index = self.obj_ids.index(n)
            rings = np.array(self.stack[index], dtype=np.float32)[None]
            m = self.m[index][None].copy()
            pos_vel[e] = oracle.ihgp_step(rings, m, self.dt_gp, self.lpf_tau, self.cx, self.cy)[0]
            self.m[index] = m[0]
        self.spin_counter += 1                                     # unregisterOldObstacle, MOT.cpp:545-584
        period = f64(5)
        if self.spin_counter > period * f64(self.frequency):
            keep = [not (f64(now) - f64(self.stack[i][self.L - 1][3]) > period) for i in range(len(self.obj_ids))]
            self.obj_ids = [v for v, k in zip(self.obj_ids, keep) if k]
            self.stack = [v for v, k in zip(self.stack, keep) if k]
            self.m = [v for v, k in zip(self.m, keep) if k]
            self.spin_counter = 0
        return np.array(this_ids, dtype=np.int32), pos_vel
