"""Synthetic KITTI-shape inputs. Seeds are part of the measurement contract:
frame i of a batch uses seed = base + i. Pure numpy; no dataset, no network."""
import numpy as np

GRID_BOUNDS = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]      # reference core/config.py:16
VOXEL_SIZE = [0.05, 0.05, 0.1]                         # core/config.py:15
MAX_VOXELS, MAX_OCCUPANCY = 20000, 5                   # core/config.py:13-14
CAR_WLH = (1.6, 3.9, 1.56)                             # core/config.py:25


def make_cloud(seed, n=16384):
    """One (n,4) f32 LiDAR-like cloud: camera-FOV wedge, 64 elevation rings, ground plane + ~20
    car-sized boxes, clipped to GRID_BOUNDS, shuffled (training shuffles: kitti_dataset.py:154)."""
    rng = np.random.default_rng(seed)
    n_ground = int(0.6 * n)
    # ground returns: range biased to near field, azimuth inside +-40 deg
    u = rng.random(n_ground)
    r = 2.0 + 68.0 * u * u
    th = np.deg2rad(rng.uniform(-40.0, 40.0, n_ground))
    # quantise elevation into 64 rings so ground points fall on arcs like a spinning LiDAR
    ring = rng.integers(0, 64, n_ground)
    elev = np.deg2rad(-24.8 + (2.0 + 24.8) * ring / 63.0)
    rr = np.where(elev < -0.01, np.minimum(r, 1.73 / np.tan(-np.minimum(elev, -0.01))), r)
    gx, gy = rr * np.cos(th), rr * np.sin(th)
    gz = -1.73 + rng.normal(0.0, 0.02, n_ground)
    ground = np.stack([gx, gy, gz], 1)
    # object returns: points on the surfaces of ~20 car-sized boxes
    n_obj = n - n_ground
    nb = 20
    cr = 5.0 + 55.0 * rng.random(nb)
    cth = np.deg2rad(rng.uniform(-35.0, 35.0, nb))
    centers = np.stack([cr * np.cos(cth), cr * np.sin(cth), np.full(nb, -1.73 + CAR_WLH[2] / 2)], 1)
    yaw = rng.uniform(-np.pi, np.pi, nb)
    which = rng.integers(0, nb, n_obj)
    local = (rng.random((n_obj, 3)) - 0.5) * np.array([CAR_WLH[1], CAR_WLH[0], CAR_WLH[2]])
    face = rng.integers(0, 3, n_obj)  # snap one coordinate to a face: surface, not volume
    sign = np.where(rng.random(n_obj) < 0.5, -0.5, 0.5)
    ext = np.array([CAR_WLH[1], CAR_WLH[0], CAR_WLH[2]])
    local[np.arange(n_obj), face] = sign * ext[face]
    c, s = np.cos(yaw[which]), np.sin(yaw[which])
    ox = centers[which, 0] + c * local[:, 0] - s * local[:, 1]
    oy = centers[which, 1] + s * local[:, 0] + c * local[:, 1]
    oz = centers[which, 2] + local[:, 2]
    obj = np.stack([ox, oy, oz], 1)
    xyz = np.concatenate([ground, obj], 0)
    lo, hi = np.array(GRID_BOUNDS[:3]), np.array(GRID_BOUNDS[3:])
    xyz = np.clip(xyz, lo + 1e-3, hi - 1e-3)
    pts = np.concatenate([xyz, rng.random((n, 1))], 1).astype(np.float32)
    rng.shuffle(pts)
    return pts


def make_batch(base_seed, batch, n=16384):
    return [make_cloud(base_seed + i, n) for i in range(batch)]


def make_nms_boxes(seed, n, group=100, degrees=False):
    """NMS microbench boxes: x~U(0,70.4) y~U(-40,40) w=1.6 l=3.9, angle~U(0,pi) radians
    exactly as the reference feeds it (a degrees-based op fed radians), scores~U(0,1); groups of
    `group` boxes separated with the wrapper's fp32 coordinate-offset trick (ops/iou_nms.py:124-132)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 70.4, n)
    y = rng.uniform(-40, 40, n)
    a = rng.uniform(0, 180.0 if degrees else np.pi, n)
    boxes = np.stack([x, y, np.full(n, 1.6), np.full(n, 3.9), a], 1).astype(np.float32)
    scores = rng.random(n).astype(np.float32)
    idxs = (np.arange(n) // group).astype(np.int64)
    return boxes, scores, idxs


def apply_group_offsets(boxes, idxs):
    """numpy restatement of the offset trick in batched_nms_rotated (ops/iou_nms.py:121-132), fp32."""
    b = boxes.astype(np.float32)
    mx = (np.maximum(b[:, 0], b[:, 1]) + np.maximum(b[:, 2], b[:, 3]) / np.float32(2)).max()
    mn = (np.minimum(b[:, 0], b[:, 1]) - np.minimum(b[:, 2], b[:, 3]) / np.float32(2)).min()
    off = idxs.astype(np.float32) * np.float32(mx - mn + np.float32(1))
    out = b.copy()
    out[:, :2] += off[:, None]
    return out


def make_active_sites(seed, n, shape, batch=1):
    """n unique active voxels per batch item sampled without replacement from a (z,y,x) grid (C4)."""
    rng = np.random.default_rng(seed)
    vol = int(shape[0]) * int(shape[1]) * int(shape[2])
    rows = []
    for b in range(batch):
        flat = rng.choice(vol, size=n, replace=False)
        z, rem = np.divmod(flat, shape[1] * shape[2])
        y, x = np.divmod(rem, shape[2])
        rows.append(np.stack([np.full(n, b), z, y, x], 1))
    return np.concatenate(rows, 0).astype(np.int32)


def make_clustered_sites(seed, n, shape, batch=1, blob=6):
    """Active sites in compact blobs (surface-like occupancy: P/N well above isolated voxels)."""
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(batch):
        got = set()
        while len(got) < n:
            c = [rng.integers(0, s) for s in shape]
            ext = [min(blob, shape[0]), blob, blob]
            for _ in range(blob * blob * 2):
                p = tuple(int(np.clip(c[d] + rng.integers(-ext[d], ext[d] + 1), 0, shape[d] - 1)) for d in range(3))
                got.add(p)
                if len(got) >= n:
                    break
        arr = np.array(sorted(got)[:n], dtype=np.int32)
        rng.shuffle(arr)
        rows.append(np.concatenate([np.full((n, 1), b, np.int32), arr], 1))
    return np.concatenate(rows, 0).astype(np.int32)
