#!/usr/bin/env python
"""Pins oracle/orb_oracle.cpp against REAL OpenCV code (the cv2 wheel in the build container) and
writes the golden ORB fixtures under tests/golden/.

Run in the build container only (needs cv2):   python oracle/pin_orb_against_cv2.py [--write]

Checks, primitive by primitive and then end to end:
  1. resize INTER_LINEAR chain        == cv2.resize
  2. reflect-101 border               == cv2.copyMakeBorder
  3. FAST-9-16 + NMS on cell ROIs     == cv2.FastFeatureDetector (thresholds 20 and 7), incl. order
  4. fastAtan2                        == cv2.fastAtan2
  5. Gaussian kernel / blur           == cv2.getGaussianKernel / cv2.sepFilter2D, and
                                         cv2.GaussianBlur applied to the non-isolated sub-matrix
  6. whole extractor                  == a Python composition of the reference's orchestration
                                         (src/ORBextractor.cpp:531-831) over those cv2 primitives
retainBest is not exposed by cv2; it is std::nth_element+std::partition in the oracle and is
cross-checked through SIFT's nfeatures path (which calls KeyPointsFilter::retainBest).
"""
import ctypes as C
import math
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tools import synth  # noqa: E402

cv2.setNumThreads(1)
lib = C.CDLL(os.path.join(HERE, "liboracle.so"))
lib.orb_oracle_create.restype = C.c_void_p
lib.orb_oracle_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int]
lib.orb_oracle_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.orb_oracle_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
lib.orb_oracle_level_dims.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
lib.orb_oracle_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.orb_oracle_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
lib.orb_oracle_fast_atan2.restype = C.c_float
lib.orb_oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
lib.orb_oracle_retain_best.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
lib.orb_oracle_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"),
                     ("octave", "i4"), ("class_id", "i4")])
PATTERN = np.loadtxt(os.path.join(HERE, "..", "se2lam_b200", "csrc", "orb_pattern_31.inc"), delimiter=",",
                     comments="//", usecols=range(32), dtype=np.int32).reshape(-1, 4)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle_extract(img, nfeatures=1000, sf=1.2, nlevels=8, fast_th=20):
    h = lib.orb_oracle_create(nfeatures, sf, nlevels, fast_th)
    kps = np.zeros(nfeatures + 64, KP_DTYPE)
    desc = np.zeros((nfeatures + 64, 32), np.uint8)
    n = lib.orb_oracle_extract(h, ptr(img), img.shape[1], img.shape[0], img.strides[0], ptr(kps), ptr(desc))
    return h, kps[:n].copy(), desc[:n].copy()


def oracle_level(h, level, blurred):
    w, hh, p = C.c_int(), C.c_int(), C.c_int()
    lib.orb_oracle_level_dims(h, level, C.byref(w), C.byref(hh), C.byref(p))
    out = np.zeros((hh.value + 32, p.value), np.uint8)
    assert lib.orb_oracle_get_level(h, level, blurred, ptr(out)) == 0
    return out, w.value, hh.value


def retain_best(resp, n):
    resp = np.ascontiguousarray(resp, np.float32)
    ids = np.zeros(len(resp), np.int32)
    m = lib.orb_oracle_retain_best(ptr(resp), len(resp), n, ptr(ids))
    return ids[:m]


# ------------------------------------------------------------------------------------------
def check_resize():
    rng = np.random.default_rng(0)
    bad = 0
    for (sw, sh, dw, dh) in [(640, 480, 533, 400), (533, 400, 444, 333), (444, 333, 370, 278), (370, 278, 309, 231),
                             (309, 231, 257, 193), (257, 193, 214, 161), (214, 161, 179, 134), (752, 480, 627, 400),
                             (100, 37, 83, 31), (17, 9, 14, 8)]:
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        out = np.zeros((dh, dw), np.uint8)
        lib.orb_oracle_resize(ptr(src), sw, sh, sw, ptr(out), dw, dh, dw)
        bad += int((ref != out).sum())
    print(f"[1] resize: {bad} differing pixels")
    return bad == 0


def check_fast():
    bad = 0
    total = 0
    for seed in (1000, 1001):
        img = synth.orb_frame(seed)
        for (x0, y0, cw, ch) in [(13, 13, 128, 81), (135, 88, 107, 68), (0, 0, 640, 480), (500, 400, 140, 80), (3, 5, 7, 7), (9, 9, 6, 20)]:
            roi = img[y0:y0 + ch, x0:x0 + cw]
            for th in (20, 7):
                det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                ref = det.detect(roi)
                ref = np.array([(k.pt[0], k.pt[1], k.response) for k in ref], np.float32).reshape(-1, 3)
                out = np.zeros((cw * ch, 3), np.float32)
                n = lib.orb_oracle_fast(ptr(img[y0:, x0:]), cw, ch, img.strides[0], th, ptr(out), cw * ch)
                total += len(ref)
                if n != len(ref) or not np.array_equal(out[:n], ref):
                    bad += 1
    print(f"[3] FAST: {bad} mismatching ROI/threshold cases over {total} corners")
    return bad == 0


def check_atan2():
    rng = np.random.default_rng(1)
    ys = rng.integers(-40000, 40000, 200000).astype(np.float32)
    xs = rng.integers(-40000, 40000, 200000).astype(np.float32)
    ys[:100] = 0
    xs[50:150] = 0
    bad = 0
    for y, x in zip(ys, xs):
        if np.float32(cv2.fastAtan2(float(y), float(x))).tobytes() != np.float32(lib.orb_oracle_fast_atan2(float(y), float(x))).tobytes():
            bad += 1
    print(f"[4] fastAtan2: {bad} mismatches / {len(ys)}")
    return bad == 0


def check_blur(h_or):
    gk = np.zeros(7, np.float32)
    dummy_i = np.zeros(16, np.int32)
    dummy_f = np.zeros(16, np.float32)
    dummy_f2 = np.zeros(16, np.float32)
    umax = np.zeros(16, np.int32)
    lib.orb_oracle_tables(h_or, ptr(dummy_i), ptr(dummy_f), ptr(dummy_f2), ptr(umax), ptr(gk))
    ref_k = cv2.getGaussianKernel(7, 2, cv2.CV_32F).reshape(-1)
    k_ok = ref_k.tobytes() == gk.tobytes()
    print(f"[5a] gaussian kernel bit-equal: {k_ok}  {gk}")
    print(f"     umax = {umax.tolist()}")
    bad_sep = bad_gb = 0
    for level in range(8):
        plain, w, hh = oracle_level(h_or, level, 0)
        blur, _, _ = oracle_level(h_or, level, 1)
        # (i) sepFilter2D on the whole bordered plane, compare the ROI interior
        ref = cv2.sepFilter2D(plain, cv2.CV_8U, ref_k, ref_k, borderType=cv2.BORDER_REFLECT_101)
        bad_sep += int((ref[16:16 + hh, 16:16 + w] != blur[16:16 + hh, 16:16 + w]).sum())
        # (ii) informational: cv2.GaussianBlur called from Python on the ROI view. numpy views carry no
        # SUBMATRIX flag, so this takes OpenCV's bit-exact fixed-point path (isolated semantics), NOT the
        # path the reference's C++ sub-matrix call takes; the difference count is printed, not asserted.
        work = plain.copy()
        roi = work[16:16 + hh, 16:16 + w]
        cv2.GaussianBlur(roi, (7, 7), 2, dst=roi, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        bad_gb += int((work != blur).sum())
    # (iii) random images, many rounding boundaries
    rng = np.random.default_rng(3)
    bad_rand = 0
    for _ in range(40):
        img = rng.integers(0, 256, (200, 300), dtype=np.uint8)
        hh2, kk, dd = oracle_extract(img, nfeatures=50, nlevels=1)
        blur, w, hh = oracle_level(hh2, 0, 1)
        plain, _, _ = oracle_level(hh2, 0, 0)
        if blur.size:
            ref = cv2.sepFilter2D(plain, cv2.CV_8U, ref_k, ref_k, borderType=cv2.BORDER_REFLECT_101)
            bad_rand += int((ref[16:16 + hh, 16:16 + w] != blur[16:16 + hh, 16:16 + w]).sum())
        lib.orb_oracle_destroy(C.c_void_p(hh2))
    print(f"[5b] blur vs cv2.sepFilter2D: {bad_sep} differing px (pyramid), {bad_rand} (40 random 300x200 images); "
          f"[info] vs Python-side cv2.GaussianBlur fixed-point path: {bad_gb} px differ")
    return k_ok and bad_sep == 0 and bad_rand == 0


# ------------------------------------------------------------------------------------------
def py_extract(img, nfeatures=1000, sf=1.2, nlevels=8, fast_th=20):
    """Independent composition of the reference orchestration over cv2 primitives."""
    f32 = np.float32
    sfd = float(f32(sf))
    mvScale = [f32(1)]
    for i in range(1, nlevels):
        mvScale.append(f32(float(mvScale[-1]) * sfd))
    inv = f32(1.0 / sfd)
    mvInv = [f32(1)]
    for i in range(1, nlevels):
        mvInv.append(f32(mvInv[-1] * inv))
    factor = f32(1.0 / sfd)
    nDes = f32(f32(nfeatures) * f32(f32(1) - factor)) / f32(f32(1) - f32(math.pow(float(factor), float(nlevels))))
    nDes = f32(nDes)
    per_level = []
    s = 0
    for lv in range(nlevels - 1):
        per_level.append(int(np.rint(nDes)))
        s += per_level[-1]
        nDes = f32(nDes * factor)
    per_level.append(max(nfeatures - s, 0))
    umax = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]

    H, W = img.shape
    planes = []
    for lv in range(nlevels):
        sw = int(np.rint(f32(f32(W) * mvInv[lv])))
        sh = int(np.rint(f32(f32(H) * mvInv[lv])))
        if lv == 0:
            cur = img
        else:
            prev = planes[lv - 1][16:-16, 16:-16]
            cur = cv2.resize(prev, (sw, sh), interpolation=cv2.INTER_LINEAR)
        planes.append(cv2.copyMakeBorder(cur, 16, 16, 16, 16, cv2.BORDER_REFLECT_101))

    ratio = f32(W) / f32(H)
    all_kps = []
    for lv in range(nlevels):
        P = planes[lv]
        roi = P[16:-16, 16:-16]
        h, w = roi.shape
        nD = per_level[lv]
        levelCols = int(np.sqrt(f32(f32(nD) / f32(f32(5) * ratio))))
        levelRows = int(f32(ratio * f32(levelCols)))
        minB = 16
        maxBX, maxBY = w - 16, h - 16
        cellW = int(math.ceil(f32(f32(maxBX - minB) / f32(levelCols))))
        cellH = int(math.ceil(f32(f32(maxBY - minB) / f32(levelRows))))
        nCells = levelRows * levelCols
        nfCell = int(math.ceil(f32(f32(nD) / f32(nCells))))
        cells = [[[] for _ in range(levelCols)] for _ in range(levelRows)]
        nToRetain = np.zeros((levelRows, levelCols), int)
        nTotal = np.zeros((levelRows, levelCols), int)
        noMore = np.zeros((levelRows, levelCols), bool)
        iniXCol = [0] * levelCols
        iniYRow = [0] * levelRows
        nNoMore = 0
        nToDist = 0
        hY = cellH + 6
        for i in range(levelRows):
            iniY = minB + i * cellH - 3
            iniYRow[i] = iniY
            if i == levelRows - 1:
                hY = maxBY + 3 - iniY
                if hY <= 0:
                    continue
            hX = cellW + 6
            for j in range(levelCols):
                if i == 0:
                    iniX = minB + j * cellW - 3
                    iniXCol[j] = iniX
                else:
                    iniX = iniXCol[j]
                if j == levelCols - 1:
                    hX = maxBX + 3 - iniX
                    if hX <= 0:
                        continue
                cell = roi[iniY:iniY + hY, iniX:iniX + hX] if iniY >= 0 and iniX >= 0 else None
                # cell ROI may start in the border ring (ini = 13 >= 0 always for EDGE 16)
                det = cv2.FastFeatureDetector_create(fast_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                k = det.detect(cell)
                if len(k) <= 3:
                    det = cv2.FastFeatureDetector_create(7, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                    k = det.detect(cell)
                k = [(kp.pt[0], kp.pt[1], kp.response) for kp in k]
                cells[i][j] = k
                nTotal[i, j] = len(k)
                if len(k) > nfCell:
                    nToRetain[i, j] = nfCell
                else:
                    nToRetain[i, j] = len(k)
                    nToDist += nfCell - len(k)
                    noMore[i, j] = True
                    nNoMore += 1
        while nToDist > 0 and nNoMore < nCells:
            nNew = int(f32(nfCell) + f32(math.ceil(f32(f32(nToDist) / f32(nCells - nNoMore)))))
            nToDist = 0
            for i in range(levelRows):
                for j in range(levelCols):
                    if not noMore[i, j]:
                        if nTotal[i, j] > nNew:
                            nToRetain[i, j] = nNew
                        else:
                            nToRetain[i, j] = nTotal[i, j]
                            nToDist += nNew - nTotal[i, j]
                            noMore[i, j] = True
                            nNoMore += 1
        size = float(int(f32(31) * mvScale[lv]))
        kps = []
        for i in range(levelRows):
            for j in range(levelCols):
                k = cells[i][j]
                if not k:
                    continue
                keep = retain_best([r for (_, _, r) in k], int(nToRetain[i, j]))
                keep = keep[:nToRetain[i, j]]
                for idx in keep:
                    x, y, r = k[idx]
                    kps.append([x + iniXCol[j], y + iniYRow[i], r])
        if len(kps) > nD:
            keep = retain_best([r for (_, _, r) in kps], nD)[:nD]
            kps = [kps[i] for i in keep]
        # orientation on the un-blurred level
        out = []
        for (x, y, r) in kps:
            cx_, cy_ = int(np.rint(x)), int(np.rint(y))
            m01 = m10 = 0
            for v in range(-15, 16):
                d = umax[abs(v)]
                row = roi[cy_ + v, cx_ - d:cx_ + d + 1].astype(np.int64)
                u = np.arange(-d, d + 1)
                m10 += int((u * row).sum())
                m01 += v * int(row.sum())
            ang = cv2.fastAtan2(float(m01), float(m10))
            out.append((x, y, size, ang, r, lv, -1))
        all_kps.append(out)

    gk = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    res_k, res_d = [], []
    for lv in range(nlevels):
        if not all_kps[lv]:
            continue
        # the reference's in-place GaussianBlur on a non-isolated sub-matrix == sepFilter2D whose
        # out-of-ROI neighbours are the real border-ring pixels; from Python (no SUBMATRIX flag on
        # numpy views) that is: filter the whole bordered plane, keep the ROI interior.
        work = planes[lv].copy()
        filt = cv2.sepFilter2D(planes[lv], cv2.CV_8U, gk, gk, borderType=cv2.BORDER_REFLECT_101)
        work[16:-16, 16:-16] = filt[16:-16, 16:-16]
        roi = work[16:-16, 16:-16]
        for (x, y, size, ang, r, octv, cid) in all_kps[lv]:
            angle = f32(f32(ang) * f32(np.pi / f32(180.0)))
            a = f32(math.cos(float(angle)))
            b = f32(math.sin(float(angle)))
            cx_, cy_ = int(np.rint(x)), int(np.rint(y))
            px0 = PATTERN[:, 0].astype(f32); py0 = PATTERN[:, 1].astype(f32)
            px1 = PATTERN[:, 2].astype(f32); py1 = PATTERN[:, 3].astype(f32)
            r0 = np.rint(px0 * b + py0 * a).astype(int); c0 = np.rint(px0 * a - py0 * b).astype(int)
            r1 = np.rint(px1 * b + py1 * a).astype(int); c1 = np.rint(px1 * a - py1 * b).astype(int)
            t0 = roi[cy_ + r0, cx_ + c0].astype(int) if False else work[16 + cy_ + r0, 16 + cx_ + c0].astype(int)
            t1 = work[16 + cy_ + r1, 16 + cx_ + c1].astype(int)
            bits = (t0 < t1).astype(np.uint8).reshape(32, 8)
            res_d.append((bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8))
            sc = mvScale[lv]
            xx, yy = (f32(f32(x) * sc), f32(f32(y) * sc)) if lv else (f32(x), f32(y))
            res_k.append((xx, yy, size, ang, r, octv, cid))
    kps = np.array(res_k, KP_DTYPE) if res_k else np.zeros(0, KP_DTYPE)
    desc = np.array(res_d, np.uint8).reshape(-1, 32)
    return kps, desc


def check_full(write):
    ok = True
    cases = [("synth1000", synth.orb_frame(1000)), ("synth1001", synth.orb_frame(1001)),
             ("constant", synth.orb_adversarial("constant")), ("noise", synth.orb_adversarial("noise")),
             ("lowcontrast", synth.orb_adversarial("lowcontrast")), ("gradient", synth.orb_adversarial("gradient")),
             ("small_320x240", synth.orb_frame(5, 320, 240)), ("odd_501x377", synth.orb_frame(6, 501, 377))]
    gold = {}
    for name, img in cases:
        h, k_or, d_or = oracle_extract(img)
        k_py, d_py = py_extract(img)
        same = len(k_or) == len(k_py) and k_or.tobytes() == k_py.tobytes() and d_or.tobytes() == d_py.tobytes()
        nd = -1
        if len(k_or) == len(k_py):
            nd = int((d_or != d_py).any(axis=1).sum())
        print(f"[6] {name}: oracle {len(k_or)} kps, cv2-composition {len(k_py)} kps, identical={same}, differing desc rows={nd}")
        ok &= same
        gold[name + "_img"] = img if name.startswith(("small", "odd")) else np.zeros(0, np.uint8)
        gold[name + "_kps"] = k_or
        gold[name + "_desc"] = d_or
        lib.orb_oracle_destroy(C.c_void_p(h))
    if write and ok:
        out = os.path.join(HERE, "..", "tests", "golden", "orb_golden.npz")
        np.savez_compressed(out, **gold)
        print("wrote", out, os.path.getsize(out), "bytes")
    return ok


def check_retain_best_via_sift():
    """cv2.SIFT(nfeatures=n) == retainBest(all, n) applied to the sorted de-duplicated list [upstream
    SIFT_Impl::detectAndCompute].  Lets us observe the real KeyPointsFilter::retainBest permutation."""
    img = synth.orb_frame(1000)
    full = cv2.SIFT_create(nfeatures=0).detect(img)
    ok = True
    for n in (50, 200, 777):
        sub = cv2.SIFT_create(nfeatures=n).detect(img)
        resp = np.array([k.response for k in full], np.float32)
        keep = retain_best(resp, n)
        a = [(full[i].pt, full[i].response) for i in keep]
        b = [(k.pt, k.response) for k in sub]
        same = a == b
        print(f"[r] retainBest via SIFT nfeatures={n}: {len(sub)} kept, oracle {len(keep)} kept, same order={same}")
        ok &= same
    return ok


if __name__ == "__main__":
    write = "--write" in sys.argv
    h, _, _ = oracle_extract(synth.orb_frame(1000))
    results = [check_resize(), check_fast(), check_atan2(), check_blur(h)]
    try:
        results.append(check_retain_best_via_sift())
    except Exception as e:  # SIFT internals are a courtesy cross-check only
        print("[r] SIFT cross-check unavailable:", e)
    results.append(check_full(write))
    print("ALL PINNED" if all(results) else "PIN FAILURES", results)
    sys.exit(0 if all(results) else 1)
