"""TEST INFRASTRUCTURE (see oracle/__init__.py): CPU restatement of JointAngleDataset.__getitem__
(/root/reference/dataset/joint_angle.py:21-233), numpy float32, statement by statement, consuming a given uniform stream in
the reference's order.  Pinned against the reference itself by tests/golden/joint_angle.npz (oracle/make_golden_poses.py)."""
import numpy as np

f32 = np.float32
PI = f32(np.pi)


def _deg(a):
    return f32(f32(a * PI) / f32(180))


class _Stream:
    def __init__(self, u, p=0):
        self.u, self.p = np.asarray(u, np.float32), p

    def rand(self):
        v = self.u[self.p]
        self.p += 1
        return f32(v)


def _curled(s, b1, b2, b3):
    """_set_closed_finger / _set_pinching_finger / _set_half_open_finger (:42-103): they differ in the additive constants."""
    def curr(base):
        a = _deg(f32(f32(s.rand() * f32(30)) + f32(base)))
        b = _deg(f32(f32(s.rand() * f32(20)) - f32(10)))
        return f32(a + b)
    f1, f2, f3 = f32(-0.2), f32(-0.4), f32(-0.34)
    c = curr(b1)
    f1 = f32(f1 + c)
    f2 = f32(f2 + f32(f32(0.2) * c))
    c = curr(b2)
    f1 = f32(f1 + f32(f32(0.2) * c))
    f2 = f32(f2 + c)
    f3 = f32(f3 + f32(f32(0.7) * c))
    c = curr(b3)
    f2 = f32(f2 + f32(f32(0.2) * c))
    f3 = f32(f3 + c)
    return [f1, f2, f3]


def _finger(s, shape):
    if shape == 0:      # _set_straight_finger :111-115
        return [f32(f32(s.rand() * f32(0.25)) - f32(0.25)), f32(f32(s.rand() * f32(0.4)) - f32(0.4)), f32(f32(s.rand() * f32(0.34)) - f32(0.34))]
    if shape == 1:      # _set_open_finger :105-109
        return [f32(f32(s.rand() * f32(0.25)) - f32(0.1)), f32(f32(s.rand() * f32(0.4)) - f32(0.1)), f32(f32(s.rand() * f32(0.34)) - f32(0.1))]
    if shape == 2:
        return _curled(s, 0, 60, 60)
    if shape == 3:
        return _curled(s, 60, 5, 5)
    return _curled(s, 60, 60, 60)


def joint_angle_getitem(s):
    """One pose [26] from the stream `s` (advances it)."""
    p = np.zeros(26, np.float32)
    p[0] = f32(f32(s.rand() * f32(6.28)) - f32(3.14))
    p[1] = f32(-s.rand() * f32(3.14))
    p[2] = f32(f32(s.rand() * f32(6.28)) - f32(3.14))
    p[3] = f32(f32(s.rand() * f32(30)) - f32(15))
    p[4] = f32(f32(s.rand() * f32(30)) - f32(15))
    p[5] = f32(f32(s.rand() * f32(50)) - f32(35))
    spread = f32(f32(s.rand() - f32(0.35)) / f32(1.55))
    for f, k in enumerate((1.55, 0.75, -0.75, -2.2)):
        r = _deg(f32(f32(s.rand() * f32(10)) - f32(5)))
        p[6 + 4 * f] = f32(f32(k) * f32(spread + r))
    sel, v = s.rand(), s.rand()
    flex = f32(f32(v * f32(0.35)) - f32(0.25)) if sel < f32(0.5) else f32(f32(v * f32(0.6)) + f32(0.1))
    f3 = f32(f32(s.rand() * f32(2)) - f32(1.7))
    p[22] = f32(s.rand() - f32(0.5))
    p[23], p[24], p[25] = flex, f32(f32(0.25) * flex), f3
    mode = int(f32(s.rand() * f32(10)))
    if mode <= 4:
        rules = (mode,) * 4
    else:
        rules = {5: (-1, -2, -2, -2), 6: (-2, -2, -2, -1), 7: (-1, -1, -2, -2), 8: (-2, -1, -1, -1)}.get(mode, (-3,) * 4)
    for f, rule in enumerate(rules):
        if rule == -1:
            shape = int(f32(s.rand() * f32(3)))
        elif rule == -2:
            shape = 3 + int(f32(s.rand() * f32(2)))
        elif rule == -3:
            shape = int(f32(s.rand() * f32(5)))
        else:
            shape = rule
        p[7 + 4 * f: 10 + 4 * f] = _finger(s, shape)
    return p


def joint_angle_batch(u, n):
    """n consecutive poses from ONE sequential stream u -> (poses [n,26], offsets [n], uniforms consumed)."""
    s = _Stream(u)
    poses, offs = [], []
    for _ in range(n):
        offs.append(s.p)
        poses.append(joint_angle_getitem(s))
    return np.stack(poses), np.asarray(offs, np.int32), s.p
