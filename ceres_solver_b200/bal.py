"""Host-side problem preparation that sits above the C ABI (what bundle_adjuster + the Ceres preprocessor do
before the hot path starts):

  read_bal / normalize      examples/bal_problem.cc:73-133, :249-292  (BAL text format, BALProblem::Normalize)
  reduced_program           the ordering the adapters read off ceres::internal::Program:
                            reorder_program.cc:217-276 (ApplyOrdering, points = group 0, first-use order inside a
                            group) and :278-359 (rows bucketed by e block, each bucket filled back to front)
  synthetic_bal             SURVEY §8d I2/I3 synthetic regeneration of Ladybug-1723 / Venice-1778-shaped problems
                            (the real files are not shipped with the reference and there is no network)
"""
import bz2

import numpy as np


class Bal:
    def __init__(self, cam_idx, pt_idx, obs, cameras, points):
        self.cam_idx = np.ascontiguousarray(cam_idx, dtype=np.int32)
        self.pt_idx = np.ascontiguousarray(pt_idx, dtype=np.int32)
        self.obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 2)
        self.cameras = np.ascontiguousarray(cameras, dtype=np.float64).reshape(-1, 9)
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)

    @property
    def C(self):
        return self.cameras.shape[0]

    @property
    def P(self):
        return self.points.shape[0]

    @property
    def N(self):
        return self.cam_idx.shape[0]


def read_bal(path):
    opener = bz2.open if path.endswith(".bz2") else open
    with opener(path, "rb") as f:
        data = np.array(f.read().split(), dtype=np.float64)
    C, P, N = int(data[0]), int(data[1]), int(data[2])
    o = data[3:3 + 4 * N].reshape(N, 4)
    params = data[3 + 4 * N:3 + 4 * N + 9 * C + 3 * P]
    if params.size != 9 * C + 3 * P:
        raise IOError("Invalid UW data file.")
    return Bal(o[:, 0].astype(np.int32), o[:, 1].astype(np.int32), o[:, 2:4].copy(), params[:9 * C], params[9 * C:])


def angle_axis_rotate(w, X):
    """include/ceres/rotation.h:864-930, vectorised over rows."""
    w = np.atleast_2d(w)
    X = np.atleast_2d(X)
    theta = np.sqrt((w * w).sum(axis=1))
    out = np.empty_like(X, dtype=np.float64)
    nz = theta != 0.0
    if nz.any():
        t = theta[nz][:, None]
        a = w[nz] / t
        Xn = X[nz]
        c, s = np.cos(t), np.sin(t)
        out[nz] = Xn * c + np.cross(a, Xn) * s + a * ((a * Xn).sum(axis=1, keepdims=True) * (1.0 - c))
    if (~nz).any():
        out[~nz] = X[~nz] + np.cross(w[~nz], X[~nz])
    return out


def _upper_median(v):
    k = v.size // 2
    return np.partition(v, k)[k]


def normalize(bal):
    """BALProblem::Normalize (bal_problem.cc:249-292): median-centre, scale so the median absolute deviation is 100."""
    pts = bal.points
    median = np.array([_upper_median(pts[:, i]) for i in range(3)])
    mad = _upper_median(np.abs(pts - median).sum(axis=1))
    scale = 100.0 / mad
    new_pts = scale * (pts - median)
    cams = bal.cameras.copy()
    w = cams[:, 0:3]
    center = -angle_axis_rotate(-w, cams[:, 3:6])          # c = -R' t
    center = scale * (center - median)
    cams[:, 3:6] = -angle_axis_rotate(w, center)           # t = -R c
    return Bal(bal.cam_idx, bal.pt_idx, bal.obs, cams, new_pts)


class ReducedProgram:
    """Row / column order of the reduced program for a BAL problem, and the state vector in that order."""

    def __init__(self, bal):
        N = bal.N
        # first-use order of parameter blocks inside each elimination group
        first_pt = np.full(bal.P, N, dtype=np.int64)
        np.minimum.at(first_pt, bal.pt_idx, np.arange(N))
        first_cam = np.full(bal.C, N, dtype=np.int64)
        np.minimum.at(first_cam, bal.cam_idx, np.arange(N))
        self.point_of_eblock = np.argsort(first_pt, kind="stable")[:int((first_pt < N).sum())].astype(np.int32)
        self.camera_of_fblock = np.argsort(first_cam, kind="stable")[:int((first_cam < N).sum())].astype(np.int32)
        e_of_point = np.full(bal.P, -1, dtype=np.int32)
        e_of_point[self.point_of_eblock] = np.arange(self.point_of_eblock.size, dtype=np.int32)
        f_of_camera = np.full(bal.C, -1, dtype=np.int32)
        f_of_camera[self.camera_of_fblock] = np.arange(self.camera_of_fblock.size, dtype=np.int32)
        e = e_of_point[bal.pt_idx]
        # rows grouped by e block; inside a block in REVERSE input order
        self.obs_of_row = np.lexsort((-np.arange(N), e)).astype(np.int32)
        self.row_pt = e[self.obs_of_row]
        self.row_cam = f_of_camera[bal.cam_idx[self.obs_of_row]]
        self.row_obs = bal.obs[self.obs_of_row]
        self.C = int(self.camera_of_fblock.size)
        self.P = int(self.point_of_eblock.size)
        self.N = N
        self.num_parameters = 3 * self.P + 9 * self.C

    def state(self, bal):
        return np.concatenate([bal.points[self.point_of_eblock].ravel(), bal.cameras[self.camera_of_fblock].ravel()])

    def shard(self, rank, world):
        """Point-range shard balanced by observation count: returns (pt_lo, pt_hi, row_lo, row_hi)."""
        counts = np.bincount(self.row_pt, minlength=self.P)
        ptr = np.concatenate([[0], np.cumsum(counts)])
        bounds = [int(np.searchsorted(ptr, self.N * r / world)) for r in range(world + 1)]
        bounds[0], bounds[-1] = 0, self.P
        lo, hi = bounds[rank], bounds[rank + 1]
        return lo, hi, int(ptr[lo]), int(ptr[hi])


def snavely_project(cameras, points, cam_idx, pt_idx):
    """examples/snavely_reprojection_error.h:57-92 predicted image point (vectorised)."""
    cam = cameras[cam_idx]
    p = angle_axis_rotate(cam[:, 0:3], points[pt_idx]) + cam[:, 3:6]
    xp = -p[:, 0] / p[:, 2]
    yp = -p[:, 1] / p[:, 2]
    r2 = xp * xp + yp * yp
    d = 1.0 + r2 * (cam[:, 7] + cam[:, 8] * r2)
    return np.stack([cam[:, 6] * d * xp, cam[:, 6] * d * yp], axis=1)


SHAPES = {
    # name: (cameras, points, observations)   public BAL listing
    "ladybug-1723": (1723, 156502, 678718),
    "venice-1778": (1778, 993923, 5001946),
    "trafalgar-257": (257, 65132, 225911),
    "tiny": (12, 300, 1500),
}


def synthetic_bal(num_cameras, num_points, num_observations, seed=38401, max_degree=64):
    """Seeded BAL-shaped problem: cameras on a circle of radius 100 looking at the origin,
    f~U(500,1500), l1~N(0,1e-7), l2~N(0,1e-13); points ~N(0,20^2)^3; each point seen by the `degree` cameras
    nearest in angle, degrees drawn to hit num_observations exactly (min 2); observations = projection +
    N(0,0.5^2) px; initial parameters = truth perturbed by rotation 1e-3, translation 1e-1, point 1e-1."""
    from scipy.spatial.transform import Rotation

    rng = np.random.RandomState(seed)  # MT19937, the example's own generator family (bundle_adjuster.cc:138)
    C, P, N = int(num_cameras), int(num_points), int(num_observations)
    max_degree = min(max_degree, C)
    if not (2 * P <= N <= max_degree * P):
        raise ValueError("need 2P <= N <= max_degree*P")
    phi = 2.0 * np.pi * np.arange(C) / C
    centers = np.stack([100.0 * np.cos(phi), 100.0 * np.sin(phi), np.zeros(C)], axis=1)
    xax = np.stack([-np.sin(phi), np.cos(phi), np.zeros(C)], axis=1)
    zax = np.stack([np.cos(phi), np.sin(phi), np.zeros(C)], axis=1)   # camera looks down its -z axis
    yax = np.cross(zax, xax)
    R = np.stack([xax, yax, zax], axis=1)
    w = Rotation.from_matrix(R).as_rotvec()
    t = -np.einsum("cij,cj->ci", R, centers)
    cams = np.zeros((C, 9))
    cams[:, 0:3] = w
    cams[:, 3:6] = t
    cams[:, 6] = rng.uniform(500.0, 1500.0, C)
    cams[:, 7] = rng.normal(0.0, 1e-7, C)
    cams[:, 8] = rng.normal(0.0, 1e-13, C)
    pts = rng.normal(0.0, 20.0, (P, 3))
    rad = np.sqrt(pts[:, 0] ** 2 + pts[:, 1] ** 2)
    pts[:, :2] *= np.minimum(1.0, 80.0 / np.maximum(rad, 1e-12))[:, None]
    # degrees: 2 + geometric tail, then exact fix-up to N
    extra = N - 2 * P
    deg = 2 + np.minimum(rng.geometric(1.0 / (1.0 + extra / P), P) - 1, max_degree - 2)
    diff = N - int(deg.sum())
    while diff != 0:
        if diff > 0:
            cand = np.flatnonzero(deg < max_degree)
            pick = cand[rng.randint(0, cand.size, min(diff, cand.size))]
            pick = np.unique(pick)
            deg[pick] += 1
        else:
            cand = np.flatnonzero(deg > 2)
            pick = cand[rng.randint(0, cand.size, min(-diff, cand.size))]
            pick = np.unique(pick)
            deg[pick] -= 1
        diff = N - int(deg.sum())
    # cameras: `degree` distinct cameras spread evenly (random phase) over a window of the circle centred on the
    # point's azimuth (+ jitter) — wide baselines, like the loops of a real capture, keep the depth observable.
    window = min(C, max(C // 8, max_degree))
    az = np.arctan2(pts[:, 1], pts[:, 0]) + rng.normal(0.0, 0.3, P)
    start = (np.round(az / (2.0 * np.pi) * C).astype(np.int64) - window // 2) % C
    phase = rng.uniform(0.0, 1.0, P)
    pt_idx = np.repeat(np.arange(P, dtype=np.int64), deg)
    ptr = np.concatenate([[0], np.cumsum(deg)])
    within = np.arange(N, dtype=np.int64) - ptr[pt_idx]
    cam_idx = (start[pt_idx] + np.floor((within + phase[pt_idx]) * (window / deg[pt_idx])).astype(np.int64)) % C
    obs = snavely_project(cams, pts, cam_idx, pt_idx) + rng.normal(0.0, 0.5, (N, 2))
    # perturbed initial guess
    cams0 = cams.copy()
    center = -angle_axis_rotate(-cams[:, 0:3], cams[:, 3:6])
    w0 = cams[:, 0:3] + rng.normal(0.0, 1e-3, (C, 3))
    cams0[:, 0:3] = w0
    cams0[:, 3:6] = -angle_axis_rotate(w0, center) + rng.normal(0.0, 1e-1, (C, 3))
    pts0 = pts + rng.normal(0.0, 1e-1, (P, 3))
    return Bal(cam_idx.astype(np.int32), pt_idx.astype(np.int32), obs, cams0, pts0)


def _look_at_cameras(centers, targets, rng):
    """Cameras at `centers` looking at `targets` (Snavely convention: the camera looks down its -z axis), with
    f~U(500,1500), l1~N(0,1e-7), l2~N(0,1e-13)."""
    from scipy.spatial.transform import Rotation
    C = centers.shape[0]
    z = centers - targets
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    up = np.tile(np.array([0.0, 0.0, 1.0]), (C, 1))
    x = np.cross(up, z)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)
    cams = np.zeros((C, 9))
    cams[:, 0:3] = Rotation.from_matrix(R).as_rotvec()
    cams[:, 3:6] = -np.einsum("cij,cj->ci", R, centers)
    cams[:, 6] = rng.uniform(500.0, 1500.0, C)
    cams[:, 7] = rng.normal(0.0, 1e-7, C)
    cams[:, 8] = rng.normal(0.0, 1e-13, C)
    return cams


def _degrees(rng, P, N, max_degree):
    """Per-point track lengths: 2 + geometric tail, fixed up to sum exactly to N."""
    if not (2 * P <= N <= max_degree * P):
        raise ValueError("need 2P <= N <= max_degree*P")
    extra = N - 2 * P
    deg = 2 + np.minimum(rng.geometric(1.0 / (1.0 + extra / P), P) - 1, max_degree - 2)
    diff = N - int(deg.sum())
    while diff != 0:
        cand = np.flatnonzero(deg < max_degree) if diff > 0 else np.flatnonzero(deg > 2)
        pick = np.unique(cand[rng.randint(0, cand.size, min(abs(diff), cand.size))])
        deg[pick] += 1 if diff > 0 else -1
        diff = N - int(deg.sum())
    return deg


def _finish(rng, cams, pts, cam_idx, pt_idx):
    """Noisy observations of the true geometry + the perturbed initial guess (SURVEY §8d sigmas)."""
    N = cam_idx.shape[0]
    obs = snavely_project(cams, pts, cam_idx, pt_idx) + rng.normal(0.0, 0.5, (N, 2))
    C, P = cams.shape[0], pts.shape[0]
    cams0 = cams.copy()
    center = -angle_axis_rotate(-cams[:, 0:3], cams[:, 3:6])
    w0 = cams[:, 0:3] + rng.normal(0.0, 1e-3, (C, 3))
    cams0[:, 0:3] = w0
    cams0[:, 3:6] = -angle_axis_rotate(w0, center) + rng.normal(0.0, 1e-1, (C, 3))
    pts0 = pts + rng.normal(0.0, 1e-1, (P, 3))
    return Bal(cam_idx.astype(np.int32), pt_idx.astype(np.int32), obs, cams0, pts0)


def synthetic_sequence(num_cameras, num_points, num_observations, seed=38401, max_degree=48):
    """Ladybug-like capture: the cameras are consecutive frames of a vehicle driving a loop (radius 300, looking
    outward and slightly ahead); every point is a track over `degree` CONSECUTIVE frames, created in frame order —
    the structure an incremental reconstruction of a video produces, and the one BAL's Ladybug sets have."""
    rng = np.random.RandomState(seed)
    C, P, N = int(num_cameras), int(num_points), int(num_observations)
    max_degree = min(max_degree, C)
    phi = 2.0 * np.pi * np.arange(C) / C
    radial = np.stack([np.cos(phi), np.sin(phi), np.zeros(C)], axis=1)
    tangent = np.stack([-np.sin(phi), np.cos(phi), np.zeros(C)], axis=1)
    centers = 300.0 * radial + np.stack([np.zeros(C), np.zeros(C), 1.5 * np.sin(7 * phi)], axis=1)
    cams = _look_at_cameras(centers, centers + 10.0 * radial + 3.0 * tangent, rng)
    deg = _degrees(rng, P, N, max_degree)
    first = np.minimum(rng.randint(0, C, P), C - deg)            # no wrap-around at the end of the sequence
    order = np.argsort(first, kind="stable")                     # points are created in frame order
    first, deg = first[order], deg[order]
    mid = first + deg // 2
    depth = rng.uniform(12.0, 40.0, P)
    pts = centers[mid] + depth[:, None] * radial[mid] + (0.2 * depth * rng.normal(0.0, 1.0, P))[:, None] * tangent[mid]
    pts[:, 2] += 0.2 * depth * rng.normal(0.0, 1.0, P)
    pt_idx = np.repeat(np.arange(P, dtype=np.int64), deg)
    ptr = np.concatenate([[0], np.cumsum(deg)])
    within = np.arange(N, dtype=np.int64) - ptr[pt_idx]
    cam_idx = first[pt_idx] + within
    return _finish(rng, cams, pts, cam_idx, pt_idx)


def synthetic_clusters(num_cameras, num_points, num_observations, seed=38401, max_degree=48, cams_per_cluster=30):
    """Venice-like photo collection: groups of cameras photograph the same facade from a wide arc; a point on a
    facade is seen by `degree` distinct cameras of its group (10 % of the tracks also reach into the next group)."""
    rng = np.random.RandomState(seed)
    C, P, N = int(num_cameras), int(num_points), int(num_observations)
    K = max(1, C // cams_per_cluster)
    cl_of_cam = np.minimum(np.arange(C) * K // C, K - 1)
    cl_start = np.searchsorted(cl_of_cam, np.arange(K))
    cl_size = np.diff(np.concatenate([cl_start, [C]]))
    max_degree = int(min(max_degree, cl_size.min()))
    ang = 2.0 * np.pi * np.arange(K) / K
    cl_center = 600.0 * np.stack([np.cos(ang), np.sin(ang), np.zeros(K)], axis=1)
    cl_normal = np.stack([np.cos(ang), np.sin(ang), np.zeros(K)], axis=1)        # facade faces outward
    cl_tangent = np.stack([-np.sin(ang), np.cos(ang), np.zeros(K)], axis=1)
    kc = cl_of_cam
    lateral = rng.uniform(-30.0, 30.0, C)
    dist = rng.uniform(35.0, 70.0, C)
    centers = cl_center[kc] + dist[:, None] * cl_normal[kc] + lateral[:, None] * cl_tangent[kc]
    centers[:, 2] += rng.uniform(-3.0, 8.0, C)
    targets = cl_center[kc] + rng.normal(0.0, 3.0, (C, 1)) * cl_tangent[kc]
    cams = _look_at_cameras(centers, targets, rng)
    deg = _degrees(rng, P, N, max_degree)
    kp = np.sort(rng.randint(0, K, P))                                            # points grouped by facade
    pts = cl_center[kp] + rng.uniform(-20.0, 20.0, P)[:, None] * cl_tangent[kp] + rng.normal(0.0, 1.5, P)[:, None] * cl_normal[kp]
    pts[:, 2] += rng.uniform(-8.0, 12.0, P)
    pt_idx = np.repeat(np.arange(P, dtype=np.int64), deg)
    ptr = np.concatenate([[0], np.cumsum(deg)])
    within = np.arange(N, dtype=np.int64) - ptr[pt_idx]
    # `degree` distinct cameras of the group: evenly spaced with a random phase, then a random rotation of the group
    size = cl_size[kp][pt_idx]
    phase = rng.uniform(0.0, 1.0, P)[pt_idx]
    rot = rng.randint(0, 1 << 20, P)[pt_idx]
    local = (np.floor((within + phase) * (size / deg[pt_idx])).astype(np.int64) + rot) % size
    cam_idx = cl_start[kp][pt_idx] + local
    return _finish(rng, cams, pts, cam_idx, pt_idx)


GENERATORS = {
    "ladybug-1723": synthetic_sequence,   # video sequence
    "trafalgar-257": synthetic_clusters,
    "venice-1778": synthetic_clusters,    # photo collection
    "tiny": synthetic_bal,
    "ladybug-1723-random": synthetic_bal,  # worst case for locality: every point sees cameras all over the index range
    "venice-1778-random": synthetic_bal,
}


def synthetic(name, seed=38401):
    C, P, N = SHAPES[name.replace("-random", "")]
    return GENERATORS[name](C, P, N, seed=seed)


def write_bal(bal, path):
    """BAL text format, as BALProblem::WriteToFile (bal_problem.cc:137-178) lays it out."""
    with open(path, "w") as f:
        f.write("%d %d %d\n" % (bal.C, bal.P, bal.N))
        for c, p, o in zip(bal.cam_idx, bal.pt_idx, bal.obs):
            f.write("%d %d %.16e %.16e\n" % (c, p, o[0], o[1]))
        for v in bal.cameras.ravel():
            f.write("%.16e\n" % v)
        for v in bal.points.ravel():
            f.write("%.16e\n" % v)
