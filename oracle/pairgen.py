"""CPU restatement of the synthetic (PD)S-COCO pair generator -- TEST INFRASTRUCTURE ONLY.

Follows (paths relative to /root/reference):

  draw_photometric / apply_photometric   src/data/transforms.py:296-330 (PhotometricDistortSimple) with
                                         :145-159 brightness, :162-176 contrast, :179-197 cvtColor,
                                         :200-211 saturation, :214-225 hue, :256-268 lighting noise
  draw_params / make_pair                src/data/transforms.py:456-576,724-725 (HomographyNetPrep.__call__),
                                         numpy twins src/data/utils.py:26-30 (cv2.getPerspectiveTransform) and
                                         :61-64 (np.linalg.inv + cv2.warpPerspective)
  to_network_input                       src/data/transforms.py:344-354 (DictToGrayscale), :369-378
                                         (DictStandardize), :728-743 (DictToTensor) + train.py:308-309 (.float())

The random draws are separated from their application so that (a) replaying a seeded
``numpy.random.RandomState`` in the reference's order reproduces the reference bit for bit and
(b) the same explicit parameters can be handed to the GPU generator (bh_pairgen_apply).
"""
import cv2
import numpy as np

PERMS = ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0))

# column layout of the per-sample parameter record shared with the CUDA generator
# (include/bihome_b200.h: bh_pair_params): two photometric blocks then geometry.
PHOTO_FIELDS = ('b_on', 'b_delta', 'contrast_first', 'c_on', 'c_alpha', 's_on', 's_alpha',
                'h_on', 'h_delta', 'l_on', 'l_perm')


def draw_photometric(rs, max_delta):
    lo = 1.0 - max_delta / 32 * 0.5
    hi = 1.0 + max_delta / 32 * 0.5
    p = {}
    p['b_on'] = int(rs.randint(2))
    p['b_delta'] = float(rs.uniform(-max_delta, max_delta)) if p['b_on'] else 0.0
    p['contrast_first'] = int(rs.randint(2))

    def contrast():
        p['c_on'] = int(rs.randint(2))
        p['c_alpha'] = float(rs.uniform(lo, hi)) if p['c_on'] else 1.0

    if p['contrast_first']:
        contrast()
    p['s_on'] = int(rs.randint(2))
    p['s_alpha'] = float(rs.uniform(lo, hi)) if p['s_on'] else 1.0
    p['h_on'] = int(rs.randint(2))
    p['h_delta'] = float(rs.uniform(-max_delta / 2, max_delta / 2)) if p['h_on'] else 0.0
    if not p['contrast_first']:
        contrast()
    p['l_on'] = 0
    p['l_perm'] = 0
    if max_delta > 0:
        p['l_on'] = int(rs.randint(2))
        if p['l_on']:
            p['l_perm'] = int(rs.randint(len(PERMS)))
    return p


def apply_photometric(image_u8, p):
    """uint8 HxWx3 RGB -> float32 HxWx3, never clipped (as the reference)."""
    im = image_u8.astype(np.float32)
    if p['b_on']:
        im += p['b_delta']
    if p['contrast_first'] and p['c_on']:
        im *= p['c_alpha']
    im = cv2.cvtColor(im, cv2.COLOR_RGB2HSV)
    if p['s_on']:
        im[:, :, 1] *= p['s_alpha']
    if p['h_on']:
        im[:, :, 0] += p['h_delta']
        im[:, :, 0][im[:, :, 0] > 360.0] -= 360.0
        im[:, :, 0][im[:, :, 0] < 0.0] += 360.0
    im = cv2.cvtColor(im, cv2.COLOR_HSV2RGB)
    if (not p['contrast_first']) and p['c_on']:
        im *= p['c_alpha']
    if p['l_on']:
        im = im[:, :, PERMS[p['l_perm']]]
    return im


def draw_params(rs, h, w, rho, patch_size, max_delta, distort=('image_1', 'image_2')):
    """Replays HomographyNetPrep's draws in the reference's order."""
    q = {'photo_1': None, 'photo_2': None}
    if 'image_1' in distort:
        q['photo_1'] = draw_photometric(rs, max_delta)
    if 'image_2' in distort:
        q['photo_2'] = draw_photometric(rs, max_delta)
    if patch_size != w:
        q['pos_x'] = int(rs.randint(rho + patch_size // 2, w - rho - patch_size // 2 + 1))
        q['pos_y'] = int(rs.randint(rho + patch_size // 2, h - rho - patch_size // 2 + 1))
    else:
        q['pos_x'], q['pos_y'] = w // 2, h // 2
    q['delta'] = rs.randint(-rho, rho, 8).reshape(4, 2)
    return q


def patch_corners(q, patch_size):
    x, y, s = q['pos_x'], q['pos_y'], patch_size // 2
    return np.array([(x - s, y - s), (x + s, y - s), (x + s, y + s), (x - s, y + s)])


def make_pair(image_u8, q, patch_size):
    """-> dict like HomographyNetPrep's return value (image_1/2 float32 HxWx3, patches PxPx3, ...)."""
    im1 = apply_photometric(image_u8, q['photo_1']) if q['photo_1'] is not None else np.copy(image_u8)
    im2 = apply_photometric(image_u8, q['photo_2']) if q['photo_2'] is not None else np.copy(image_u8)
    c = patch_corners(q, patch_size)
    patch_1 = im1[c[0, 1]:c[3, 1], c[0, 0]:c[1, 0]]
    H = cv2.getPerspectiveTransform(np.float32(c), np.float32(c + q['delta']))
    im2w = cv2.warpPerspective(im2, np.linalg.inv(H), dsize=(im2.shape[1], im2.shape[0]))
    patch_2 = im2w[c[0, 1]:c[3, 1], c[0, 0]:c[1, 0]]
    return {'image_1': im1, 'image_2': im2w, 'patch_1': patch_1, 'patch_2': patch_2, 'corners': c,
            'target': q['delta'], 'delta': q['delta'], 'homography': H}


def to_network_input(patch_rgb_f32, mean=0.443, std=0.129):
    """float32 PxPx3 -> float32 1xPxP as it reaches the model (fp64 standardise, then .float())."""
    g = patch_rgb_f32[:, :, 0] * 0.299 + patch_rgb_f32[:, :, 1] * 0.587 + patch_rgb_f32[:, :, 2] * 0.114
    g = np.expand_dims(g, -1)
    g = (g.astype(np.float32) / 255 - np.array([mean])) / np.array([std])      # list mean/std => float64
    return g.transpose(2, 0, 1).astype(np.float32)


def synthetic_image(index, h=240, w=320):
    """Deterministic COCO-like uint8 RGB image (smooth blobs + ramp + texture); no files needed."""
    rs = np.random.RandomState(100003 + 7919 * int(index))
    base = rs.uniform(0, 255, size=(h // 8 + 2, w // 8 + 2, 3)).astype(np.float32)
    im = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    ramp = rs.uniform(-0.25, 0.25, size=(2, 3)).astype(np.float32)
    im = im + xx[..., None] * ramp[0] + yy[..., None] * ramp[1]
    im = im + rs.normal(0, 6.0, size=(h, w, 3)).astype(np.float32)
    return np.clip(np.rint(im), 0, 255).astype(np.uint8)


def pack_params(qs):
    """list of draw dicts -> float64 [B, 2*11 + 2 + 8] table (the layout bh_pairgen_apply reads)."""
    rows = []
    for q in qs:
        r = []
        for k in ('photo_1', 'photo_2'):
            p = q[k] or dict(b_on=0, b_delta=0, contrast_first=1, c_on=0, c_alpha=1, s_on=0, s_alpha=1,
                             h_on=0, h_delta=0, l_on=0, l_perm=0)
            r += [float(p[f]) for f in PHOTO_FIELDS]
        r += [float(q['pos_x']), float(q['pos_y'])] + [float(v) for v in q['delta'].reshape(-1)]
        rows.append(r)
    return np.asarray(rows, dtype=np.float64)


def rounding_ties(q, patch_size, tol=1e-6):
    """Patch pixels whose cv2.warpPerspective source coordinate sits on a 1/32-px rounding tie.

    Integer corners and integer offsets make 32*u exactly half-integral at a handful of pixels; which way cv2 rounds
    there depends on the last bit of the (twice inverted) homography, i.e. on the LAPACK build.  Parity checks
    compare everything but these pixels.
    """
    c = patch_corners(q, patch_size)
    H = cv2.getPerspectiveTransform(np.float32(c), np.float32(c + q['delta']))
    ys, xs = np.mgrid[c[0, 1]:c[3, 1], c[0, 0]:c[1, 0]].astype(np.float64)
    w = H[2, 0] * xs + H[2, 1] * ys + H[2, 2]
    fx = 32.0 * (H[0, 0] * xs + H[0, 1] * ys + H[0, 2]) / w
    fy = 32.0 * (H[1, 0] * xs + H[1, 1] * ys + H[1, 2]) / w
    tie = lambda f: np.abs(f - np.floor(f) - 0.5) < tol
    return tie(fx) | tie(fy)
