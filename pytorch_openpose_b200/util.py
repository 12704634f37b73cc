"""Host helpers the reference's callers import from src/util.py (padRightDownCorner :12-32, transfer :36-40,
handDetect :133-201, npmax :205-210).  Drawing helpers are visualisation only and out of scope."""
import math

import numpy as np


def padRightDownCorner(img, stride, padValue):
    """Pad the bottom / right edges so both sides become multiples of `stride`.  Returns (padded, pad) with
    pad = [up, left, down, right] (up and left are always 0).  Like the reference, which builds the pad rows from the
    slice `[-2:-1]` (src/util.py:27,29), an image with a single row (column) gets NO bottom (right) padding although
    `pad` still reports it."""
    h, w = img.shape[:2]
    pad = [0, 0, (stride - h % stride) % stride, (stride - w % stride) % stride]
    down = pad[2] if h > 1 else 0
    right = pad[3] if w > 1 else 0
    out = np.full((h + down, w + right) + img.shape[2:], padValue, dtype=img.dtype)
    out[:h, :w] = img
    return out, pad


def transfer(model, model_weights):
    """Map a caffe-keyed flat checkpoint onto a module's state-dict names (drop the block prefix)."""
    return {name: model_weights[name.split(".", 1)[1]] for name in model.state_dict().keys()}


def handDetect(candidate, subset, oriImg):
    """Square hand boxes [x, y, w, is_left] from shoulder / elbow / wrist key points of every person."""
    img_h, img_w = oriImg.shape[0:2]
    boxes = []
    for person in subset.astype(int):
        for (shoulder, elbow, wrist), is_left in (((5, 6, 7), True), ((2, 3, 4), False)):
            ids = person[[shoulder, elbow, wrist]]
            if (ids == -1).any():
                continue
            (x1, y1), (x2, y2), (x3, y3) = (candidate[i][:2] for i in ids)
            reach = math.sqrt((x3 - x2) ** 2 + (y3 - y2) ** 2)
            upper = math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)
            width = 1.5 * max(reach, 0.9 * upper)
            x = x3 + 0.33 * (x3 - x2) - width / 2
            y = y3 + 0.33 * (y3 - y2) - width / 2
            if x < 0:
                x = 0
            if y < 0:
                y = 0
            w_fit = img_w - x if x + width > img_w else width
            h_fit = img_h - y if y + width > img_h else width
            boxes.append([int(x), int(y), int(min(w_fit, h_fit)), is_left])
    return boxes


def npmax(array):
    """(row, col) of the first maximum in row-major order."""
    i, j = divmod(int(np.argmax(array)), array.shape[1])
    return i, j
