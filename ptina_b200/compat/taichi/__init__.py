"""`taichi` as a PTina DRIVER SCRIPT sees it (exams/benchmark.py: `ti.init(ti.cuda)`, `ti.imshow(img, title)`; exams/matball.py:
`ti.imresize`).  ptina_b200 has no Taichi kernels: the renderer is libptina_b200.so, so there is nothing for `ti.init` to do but
record the request.  Not a Taichi implementation -- `ti.kernel`, fields and the rest of the language are absent on purpose."""
import math
import os

import numpy as np

__version__ = (0, 7, 0)
pi, tau = math.pi, math.tau


class _Arch(str):
    pass


cuda, gpu, cpu, x64, opengl, metal, cc, vulkan = (_Arch(n) for n in ('cuda', 'gpu', 'cpu', 'x64', 'opengl', 'metal', 'cc', 'vulkan'))
f32, f64, i32, i64, u8, u32 = np.float32, np.float64, np.int32, np.int64, np.uint8, np.uint32
init_args = {}
shown = []          # (title, shape) of every ti.imshow call, for scripts run headless


def init(arch=None, **kwargs):
    """Whatever backend the script asks for, rendering happens on the CUDA device of the ptina_b200 context."""
    init_args.clear()
    init_args.update(arch=arch, **kwargs)


def imshow(img, title='imshow'):
    """Headless: remembers the call; writes the image when PTINA_B200_IMSHOW_DIR is set (one .npy per call)."""
    img = np.asarray(img)
    shown.append((title, img.shape))
    out = os.environ.get('PTINA_B200_IMSHOW_DIR')
    if out:
        os.makedirs(out, exist_ok=True)
        np.save(os.path.join(out, f'imshow_{len(shown):03d}.npy'), img)


def imresize(img, w, h=None):
    """Nearest-neighbour resize of an [x, y, c] image to [w, h, c] (what the scripts use to blow up a preview)."""
    img = np.asarray(img)
    h = w if h is None else h
    xi = (np.arange(w) * img.shape[0] // w).clip(0, img.shape[0] - 1)
    yi = (np.arange(h) * img.shape[1] // h).clip(0, img.shape[1] - 1)
    return img[xi][:, yi]


def imwrite(img, path):
    """Binary PPM / raw .npy (no imaging dependency): img is [x, y, c] floats in [0, 1] with y up, like the film."""
    img = np.asarray(img)
    if path.endswith('.npy'):
        np.save(path, img)
        return
    rgb = (np.clip(img[..., :3], 0, 1) * 255 + 0.5).astype(np.uint8)
    rgb = np.ascontiguousarray(np.swapaxes(rgb, 0, 1)[::-1])
    with open(path, 'wb') as fh:
        fh.write(b'P6\n%d %d\n255\n' % (rgb.shape[1], rgb.shape[0]))
        fh.write(rgb.tobytes())
