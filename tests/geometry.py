"""Independent geometric check for the obstacle-aware controller: exact distance between the vehicle / trailer
rectangles (truck_trailer_model.py:31-72) and axis-aligned obstacle rectangles, vectorised in numpy.  The OBCA rows of
mpc_control_obs.py:65-139 are a dual certificate of ``distance >= d_min``; this module measures the distance directly."""
import numpy as np


def body_corners(states, L1=7.05, L2=12.45, M=0.15, W1=3.05, W2=2.95):
    """states [...,6] -> vehicle corners [...,4,2], trailer corners [...,4,2] (counter-clockwise)."""
    x, y, th, psi = states[..., 0], states[..., 1], states[..., 2], states[..., 3]

    def rect(cx, cy, al, L, W):
        c, s = np.cos(al), np.sin(al)
        loc = np.array([[L / 2, W / 2], [-L / 2, W / 2], [-L / 2, -W / 2], [L / 2, -W / 2]])
        px = cx[..., None] + c[..., None] * loc[:, 0] - s[..., None] * loc[:, 1]
        py = cy[..., None] + s[..., None] * loc[:, 0] + c[..., None] * loc[:, 1]
        return np.stack([px, py], axis=-1)

    veh = rect(x + np.cos(th) * L1 / 2, y + np.sin(th) * L1 / 2, th, L1, W1)
    al = th + psi
    trl = rect(x - np.cos(th) * M - np.cos(al) * L2 / 2, y - np.sin(th) * M - np.sin(al) * L2 / 2, al, L2, W2)
    return veh, trl


def box_corners(rect):
    cx, cy, w, h = rect
    return np.array([[cx + w / 2, cy + h / 2], [cx - w / 2, cy + h / 2], [cx - w / 2, cy - h / 2], [cx + w / 2, cy - h / 2]])


def _pt_seg(p, a, b):
    """distance from points p [...,P,1,2] to segments a,b [...,1,S,2]"""
    ab = b - a
    t = np.clip(((p - a) * ab).sum(-1) / (ab * ab).sum(-1), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[..., None] * ab), axis=-1)


def _inside(p, poly):
    """points p [...,P,2] strictly inside convex CCW polygon poly [...,4,2] -> [...,P]"""
    a = poly[..., None, :, :]
    b = np.roll(poly, -1, axis=-2)[..., None, :, :]
    pp = p[..., :, None, :]
    cr = (b[..., 0] - a[..., 0]) * (pp[..., 1] - a[..., 1]) - (b[..., 1] - a[..., 1]) * (pp[..., 0] - a[..., 0])
    return (cr > 0).all(-1)


def poly_distance(A, B):
    """Distance between convex CCW quadrilaterals A [...,4,2] and B [...,4,2]; negative when they overlap."""
    def vs(P, Q):
        a = Q[..., None, :, :]
        b = np.roll(Q, -1, axis=-2)[..., None, :, :]
        return _pt_seg(P[..., :, None, :], a, b).min((-1, -2))

    d = np.minimum(vs(A, B), vs(B, A))
    # overlap test by separating axes (edge normals of both)
    sep = np.zeros(d.shape, bool)
    for P in (A, B):
        e = np.roll(P, -1, axis=-2) - P
        n = np.stack([e[..., 1], -e[..., 0]], axis=-1)  # outward normals of a CCW polygon
        for i in range(4):
            ni = n[..., i, :]
            pa = (A * ni[..., None, :]).sum(-1)
            pb = (B * ni[..., None, :]).sum(-1)
            sep |= (pa.max(-1) < pb.min(-1)) | (pb.max(-1) < pa.min(-1))
    return np.where(sep, d, -d)


def clearance(states, rects, **geom):
    """min over bodies and obstacles of the body-obstacle distance, per state: states [...,6] -> [...]"""
    veh, trl = body_corners(states, **geom)
    out = np.full(states.shape[:-1], np.inf)
    for r in rects:
        box = np.broadcast_to(box_corners(r), veh.shape)
        out = np.minimum(out, np.minimum(poly_distance(veh, box), poly_distance(trl, box)))
    return out
