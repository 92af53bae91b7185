"""GridPatchSampler with the reference surface (models/sampler.py:8-354), rebuilt around integer index math.

Behavioural contract kept from the reference (checked bit-exactly against it in tests/test_sampler_parity.py):
  * numpy RNG consumption order: one np.random.uniform per sample_patches call, then one
    np.random.choice(pool, N_samples, replace=False) for the fake-patch centroids (sampler.py:260-261,324),
    plus one np.random.choice over the valid unfold patches when no_reg_sampling is set (sampler.py:229);
  * candidate real centroids = fake centroid + i*shift1 + j*shift2, i,j in [-10,10), top-1 proposal only,
    (x,y)->(row,col) swapped (sampler.py:32-35,89-94,149-153), filtered by image bounds and by the unknown-pixel
    ratio, ranked by |i|+|j| with torch.topk on the same tensor the reference builds (same tie-breaking);
  * crops are extract_glimpse(mode='nearest', normalized=False, centered=False), which for integer centroids is
    img[c-ps/2 : c+ps/2] with zero padding (utils/extract_glimpse.py:7-79).

What changed: the reference tiles the full image once per candidate (up to N_samples*400 copies,
sampler.py:171-178) and materialises every stride-ps/10 unfold patch (sampler.py:66-84).  Here the unknown-pixel
count of any window comes from a summed-area table in O(1) and only the patches that are returned are cropped.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn.functional as F  # noqa: F401  (re-exported like the reference module)


_device_gen = {}


def _mode():
    return os.environ.get("NPP_B200_SAMPLER", "parity")


def _uniform(device):
    """The patch-source draw of sample_patches (np.random.uniform(0, 1), sampler.py:324)."""
    if _mode() == "device" and device.type == "cuda":
        return float(torch.rand(1, device=device, generator=_generator(device)).item())
    return np.random.uniform(0, 1)


def _generator(device):
    g = _device_gen.get(device)
    if g is None:
        g = _device_gen[device] = torch.Generator(device=device)       # Philox, seeded from torch's global seed
        g.manual_seed(torch.initial_seed())
    return g


def _choice(n, k, device=None):
    """np.random.choice(n, size=[k], replace=False) of the reference (sampler.py:229,260).  numpy's legacy RandomState
    shuffles all n entries to draw k of them (1.9 ms for a 512 x 512 pool, more than the rest of sample_patches); the
    default keeps that call so the host RNG stream stays the reference's.  NPP_B200_SAMPLER=fast draws k distinct
    indices by rejection instead: same distribution (uniform without replacement), O(k), a different RNG stream;
    NPP_B200_SAMPLER=device does the same with the CUDA (Philox) generator, no host RNG at all."""
    if _mode() == "device" and device is not None and device.type == "cuda" and 4 * k <= n:
        g = _generator(device)
        idx = torch.randint(0, n, (k,), device=device, generator=g)
        while len(torch.unique(idx)) < k:
            idx = torch.randint(0, n, (k,), device=device, generator=g)
        return idx.cpu().numpy()
    if _mode() not in ("fast", "device") or 4 * k > n:
        return np.random.choice(n, size=[k], replace=False)
    idx = np.random.randint(0, n, size=k)
    while len(np.unique(idx)) < k:
        idx = np.random.randint(0, n, size=k)
    return idx


def _nearest_index(center, size, dim):
    """Pixel index read by sample i of a glimpse of `size` centred at `center` (float pixels) along an axis of
    length `dim`: the fp32 arithmetic of utils/extract_glimpse.py:62-79 followed by grid_sample's
    unnormalise + nearbyint (align_corners=False, mode='nearest').  Returns [M,size] int64 (may be out of range)."""
    c = center.to(torch.float32)
    k = torch.arange(0, size, dtype=torch.float32, device=c.device) - (size - 1) / 2.0
    x = c[:, None] + k[None, :]
    half = c.new_tensor(dim / 2)
    g = (x - half) / half
    ix = ((g + 1) * dim - 1) / 2
    return torch.round(ix).long()                 # torch.round is round-half-even, like nearbyint


def _is_integral(t):
    return (not t.is_floating_point()) or bool((t == torch.floor(t)).all())


def _gather_window(img_nchw, rows, cols):
    """img [1,C,H,W]; rows [M,h], cols [M,w] int64 index tables -> [M,C,h,w], zeros outside the image.
    CUDA images held as [1,H,W,C] fp32 (how the train scripts pass them) are cropped by one kernel
    (npp_gather_windows); anything else by the equivalent torch indexing."""
    _, C, H, W = img_nchw.shape
    if img_nchw.is_cuda and img_nchw.dtype == torch.float32 and rows.shape[0] > 0 and \
            img_nchw.permute(0, 2, 3, 1).is_contiguous():
        from ._core import native as _nat
        rows = rows.to(img_nchw.device, torch.int64).contiguous()
        cols = cols.to(img_nchw.device, torch.int64).contiguous()
        out = torch.empty(rows.shape[0], C, rows.shape[1], cols.shape[1], device=img_nchw.device)
        _nat.check(_nat.lib().npp_gather_windows(img_nchw.data_ptr(), H, W, C, rows.data_ptr(), cols.data_ptr(),
                                                 rows.shape[0], rows.shape[1], cols.shape[1], out.data_ptr(),
                                                 _nat.current_stream()))
        return out
    ok = ((rows >= 0) & (rows < H))[:, :, None] & ((cols >= 0) & (cols < W))[:, None, :]
    out = img_nchw[0][:, rows.clamp(0, H - 1)[:, :, None], cols.clamp(0, W - 1)[:, None, :]]   # [C,M,h,w]
    return (out * ok[None].to(out.dtype)).permute(1, 0, 2, 3).contiguous()


def extract_glimpse(input, size, offsets, centered=False, normalized=False, mode='nearest', padding_mode='zeros'):
    """Nearest-neighbour glimpse for pixel offsets (x, y) = window centre; the only configuration the sampler uses
    (utils/extract_glimpse.py with mode='nearest', normalized=False, centered=False, padding_mode='zeros')."""
    if centered or normalized or mode != 'nearest' or padding_mode != 'zeros':
        raise NotImplementedError("only the sampler's configuration is implemented")
    h, w = size
    if input.shape[0] not in (1, offsets.shape[0]):
        raise ValueError("batch size must be 1 or match the number of offsets")
    W, H = input.size(-1), input.size(-2)
    return _gather_window(input[:1], _nearest_index(offsets[:, 1], h, H), _nearest_index(offsets[:, 0], w, W))


def _crop(img_nchw, r0, c0, h, w):
    """img [1,C,H,W]; windows with top-left (r0[i], c0[i]) -> [M,C,h,w], zero padded outside the image."""
    dev = img_nchw.device
    rows = r0.to(dev)[:, None] + torch.arange(h, device=dev)[None, :]          # [M,h]
    cols = c0.to(dev)[:, None] + torch.arange(w, device=dev)[None, :]          # [M,w]
    return _gather_window(img_nchw, rows, cols)


class _UnknownCounter:
    """Summed-area table of (mask < 0.5); pixels outside the image count as unknown (zero padding)."""

    def __init__(self, mask_hw):
        unk = (mask_hw < 0.5).to(torch.int64)
        self.H, self.W = unk.shape
        sat = torch.zeros(self.H + 1, self.W + 1, dtype=torch.int64, device=unk.device)
        sat[1:, 1:] = unk.cumsum(0).cumsum(1)
        self.sat = sat

    def count(self, r0, c0, h, w):
        ra, rb = r0.clamp(0, self.H), (r0 + h).clamp(0, self.H)
        ca, cb = c0.clamp(0, self.W), (c0 + w).clamp(0, self.W)
        inside = self.sat[rb, cb] - self.sat[ra, cb] - self.sat[rb, ca] + self.sat[ra, ca]
        area = (rb - ra) * (cb - ca)
        return inside + (h * w - area)


class GridPatchSampler():
    def __init__(self, img, mask, N_samples, patch_size, height, width, pool_train, pool_val, selected_shifts,
                 no_reg_sampling):
        self.N_samples = int(N_samples)
        self.height, self.width = height, width
        self.device = img.device
        self.img, self.mask = img.permute(0, 3, 1, 2), mask.permute(0, 3, 1, 2)
        self._mask_counter = _UnknownCounter(self.mask[0, 0])
        # only the top-1 periodicity is used for sampling; first coordinate along the vertical direction
        selected_shifts = selected_shifts[0]
        self.selected_shifts = [torch.tensor([s[1], s[0]]) for s in selected_shifts]
        self._shifts_integral = all(_is_integral(s) for s in self.selected_shifts)
        r = np.arange(-10, 10)
        self._dist400 = (np.abs(r)[:, None] + np.abs(r)[None, :]).reshape(-1).astype(np.int64)   # |i| + |j|, (i, j) order
        self.no_reg_sampling = no_reg_sampling
        self.coord_patches = None
        self.reset_patchsize(img, mask, patch_size, N_samples)
        self.reset_pool(pool_train, pool_val)

    # ------------------------------------------------------------------------------------------
    def reset_patchsize(self, img, mask, patch_size, N_samples, ratio=0.0):
        self.N_samples = N_samples
        self.patch_size_h_half, self.patch_size_w_half = patch_size // 2, patch_size // 2
        self._patch_size = patch_size
        # valid positions of the stride-(patch_size//10) unfold grid (random strategy, sampler.py:66-84)
        self._unfold_img = img.permute(0, 3, 1, 2)
        self._unfold_mask = mask.permute(0, 3, 1, 2)
        stride = max(patch_size // 10, 1)
        H, W = mask.shape[1], mask.shape[2]
        ys = torch.arange(0, H - patch_size + 1, stride, device=self.device)
        xs = torch.arange(0, W - patch_size + 1, stride, device=self.device)
        if len(ys) and len(xs):
            yy, xx = torch.meshgrid(ys, xs, indexing="ij")
            cnt = _UnknownCounter(self._unfold_mask[0, 0]).count(yy.reshape(-1), xx.reshape(-1), patch_size, patch_size)
            keep = ~(cnt > (patch_size ** 2 * ratio))
            self._unfold_r0, self._unfold_c0 = yy.reshape(-1)[keep], xx.reshape(-1)[keep]
        else:
            self._unfold_r0 = self._unfold_c0 = torch.zeros(0, dtype=torch.int64, device=self.device)
        self.coord_patches = True

        self.max_shifting_ind = 10
        r = torch.arange(-self.max_shifting_ind, self.max_shifting_ind, device=self.device)
        self.permutation1, self.permutation2 = torch.meshgrid(r, r, indexing="ij")
        self.permute_distance = (abs(self.permutation1) + abs(self.permutation2)).reshape(-1).tile(self.N_samples)
        n_perm = (2 * self.max_shifting_ind) ** 2
        self.coord_batch_indicator = torch.arange(self.N_samples, device=self.device).repeat_interleave(n_perm)

    def reset_pool(self, pool_train, pool_val):
        def _get_valid_centroid(pool):
            pool = pool.to(self.device)
            valid = (pool[:, 0] > self.patch_size_h_half) & (pool[:, 0] < self.height - (self.patch_size_h_half + 1)) & \
                    (pool[:, 1] > self.patch_size_w_half) & (pool[:, 1] < self.width - (self.patch_size_w_half + 1))
            return pool[valid]

        self.pool_train = _get_valid_centroid(pool_train)
        self.pool_val = _get_valid_centroid(pool_val)
        # CUDA fast path: integer pools (the scripts build them with np.nonzero) are mirrored on the host once, so that a
        # draw's centroids, window index tables and lattice candidates need no device round trip per call
        self._pool_host = {}
        if self.device.type == "cuda":
            for name, pool in (("train", self.pool_train), ("val", self.pool_val)):
                if pool.shape[0] and _is_integral(pool):
                    self._pool_host[name] = pool.cpu().numpy().astype(np.int64)
        self._last_cent_host = None

    # ------------------------------------------------------------------------------------------
    def _sample_real_fused(self, cent, topk, invalid_ratio):
        """CUDA path of the periodicity-guided strategy for an integer lattice: ONE kernel evaluates all N x 400 candidate
        centroids (bounds + unknown-pixel ratio from the summed-area table), the tiny ranking runs on the host with the
        reference's own torch.topk call on a CPU tensor -- many candidates tie on |i| + |j| and torch.topk's tie order is
        implementation defined, so this reproduces the reference's CPU results (the goldens) on any device -- and two
        gather kernels cut the selected windows.  Replaces ~40 small CUDA launches and several host round trips."""
        from ._core import native as _nat
        hh, wh = self.patch_size_h_half, self.patch_size_w_half
        N = self.N_samples
        cent_h = cent if isinstance(cent, np.ndarray) else cent.cpu().numpy().astype(np.int64)   # [N, 2]
        cent_d = torch.as_tensor(cent_h, device=self.device)
        s1, s2 = (np.asarray(x.cpu().numpy(), np.int64) for x in self.selected_shifts)
        shifts4 = (C.c_int64 * 4)(int(s1[0]), int(s1[1]), int(s2[0]), int(s2[1]))
        keep = torch.empty(N * 400, dtype=torch.uint8, device=self.device)
        thresh = float(np.float32(hh * wh * 4 * invalid_ratio))        # torch compares an int64 tensor with a float in fp32
        _nat.check(_nat.lib().npp_sampler_candidates(self._mask_counter.sat.data_ptr(), self.height, self.width,
                                                     cent_d.data_ptr(), N, shifts4, hh, wh, thresh, keep.data_ptr(),
                                                     _nat.current_stream()))
        keep_h = keep.cpu().numpy().reshape(N, 400).astype(bool)       # the one synchronising copy
        dist_all = self._dist400
        sel_cent, weight_topks = [], []
        self.last_topk_distance = []
        topk_min = topk
        for i in range(N):
            q = np.nonzero(keep_h[i])[0]
            distance = torch.from_numpy(dist_all[q].copy())
            distance[distance == 0] = 10000                            # the patch itself is never its own reference
            if min(len(distance) - 1, topk) < topk_min:
                topk_min = min(len(distance) - 1, topk)
                if topk_min <= 0:
                    return None, None, None, 0
            distance_topk, inds_topk = torch.topk(distance, k=topk_min, largest=False)
            self.last_topk_distance.append(distance_topk.clone())
            distance_topk = 1 / distance_topk
            weight_topks.append(distance_topk / torch.sum(distance_topk))
            qs = q[inds_topk.numpy()]
            ii, jj = qs // 20 - 10, qs % 20 - 10
            sel_cent.append(cent_h[i][None, :] + ii[:, None] * s1[None, :] + jj[:, None] * s2[None, :])
        if topk_min < topk:
            weight_topks = [w[:topk_min] for w in weight_topks]
            sel_cent = [c[:topk_min] for c in sel_cent]
        weight_topks = torch.cat(weight_topks).to(self.device)
        cents = np.concatenate(sel_cent)                                # [N * topk_min, 2]
        rows = torch.as_tensor((cents[:, 0] - hh)[:, None] + np.arange(2 * hh)[None, :], device=self.device)
        cols = torch.as_tensor((cents[:, 1] - wh)[:, None] + np.arange(2 * wh)[None, :], device=self.device)
        img_p = _gather_window(self.img, rows, cols).reshape(N, topk_min, 3, 2 * hh, 2 * wh)
        mask_p = _gather_window(self.mask, rows, cols).reshape(N, topk_min, 1, 2 * hh, 2 * wh)
        return img_p.permute(0, 1, 3, 4, 2), mask_p.permute(0, 1, 3, 4, 2), weight_topks, topk_min

    def sample_patch_real(self, fake_coords=None, topk=5, invalid_ratio=0.3):
        hh, wh = self.patch_size_h_half, self.patch_size_w_half
        if fake_coords is not None and not self.no_reg_sampling and self.device.type == "cuda" and \
                self.img.dtype == torch.float32 and self._shifts_integral and fake_coords.shape[0] == self.N_samples and \
                os.environ.get("NPP_B200_SAMPLER_FUSED", "1") != "0":
            if self._last_cent_host is not None:        # the centroids sample_patch_fake just drew (host mirror)
                return self._sample_real_fused(self._last_cent_host, topk, invalid_ratio)
            cent = fake_coords[:, hh, wh, :]
            if _is_integral(cent):
                return self._sample_real_fused(cent, topk, invalid_ratio)
        if fake_coords is not None and not self.no_reg_sampling:
            N = fake_coords.shape[0]
            cent = fake_coords[:, hh, wh, :].to(self.device)                                  # [N,2] (row, col)
            s1, s2 = (s.to(self.device) for s in self.selected_shifts)
            total = s1[None, None, :] * self.permutation1[..., None] + s2[None, None, :] * self.permutation2[..., None]
            pool = (cent[:, None, None, :] + total[None]).reshape(-1, 2)                       # (sample, i, j) order
            in_bound = (pool[:, 0] > 0) & (pool[:, 0] < self.height - 1) & (pool[:, 1] > 0) & (pool[:, 1] < self.width - 1)
            sel = torch.nonzero(in_bound).squeeze(1)        # one host round trip for the three selections below
            pool = pool[sel]
            indicator = self.coord_batch_indicator[:N * total.shape[0] * total.shape[1]][sel]
            distance_all = self.permute_distance[:N * total.shape[0] * total.shape[1]][sel]
            integral = _is_integral(pool)
            if integral:      # windows are contiguous: O(1) unknown-pixel counts from the summed-area table
                r0 = pool[:, 0].long() - hh
                c0 = pool[:, 1].long() - wh
                unknown = self._mask_counter.count(r0, c0, 2 * hh, 2 * wh)
            else:             # fractional shifts: reproduce grid_sample's per-sample nearest rounding exactly
                rows = _nearest_index(pool[:, 0], 2 * hh, self.height)
                cols = _nearest_index(pool[:, 1], 2 * wh, self.width)
                unknown = (_gather_window(self.mask, rows, cols) < 0.5).sum(dim=[1, 2, 3])
            keep = torch.nonzero(~(unknown > (hh * wh * 4 * invalid_ratio))).squeeze(1)
            pool, indicator, distance_all = pool[keep], indicator[keep], distance_all[keep]

            sel_cent, weight_topks = [], []
            self.last_topk_distance = []
            topk_min = topk
            for i in range(self.N_samples):
                inds = indicator == i
                distance = distance_all[inds]
                distance[distance == 0] = 10000            # the patch itself is never its own reference
                if min(len(distance) - 1, topk) < topk_min:
                    topk_min = min(len(distance) - 1, topk)
                    if topk_min <= 0:
                        return None, None, None, 0
                # the reference's own call on the reference's own tensor: many candidates tie on |i|+|j| and torch.topk's
                # tie order is implementation defined (it differs between CPU and CUDA), so only the identical call
                # reproduces the reference on a given device
                distance_topk, inds_topk = torch.topk(distance, k=topk_min, largest=False)
                self.last_topk_distance.append(distance_topk.clone())
                distance_topk = 1 / distance_topk
                weight_topks.append(distance_topk / torch.sum(distance_topk))
                sel_cent.append(pool[inds][inds_topk])
            if topk_min < topk:
                weight_topks = [w[:topk_min] for w in weight_topks]
                sel_cent = [c[:topk_min] for c in sel_cent]
            weight_topks = torch.cat(weight_topks)
            cents = torch.cat(sel_cent)
            if integral:
                rows = (cents[:, 0].long() - hh)[:, None] + torch.arange(2 * hh, device=self.device)[None, :]
                cols = (cents[:, 1].long() - wh)[:, None] + torch.arange(2 * wh, device=self.device)[None, :]
            else:
                rows = _nearest_index(cents[:, 0], 2 * hh, self.height)
                cols = _nearest_index(cents[:, 1], 2 * wh, self.width)
            select_img_patches = _gather_window(self.img, rows, cols).reshape(self.N_samples, topk_min, 3, 2 * hh, 2 * wh)
            select_mask_patches = _gather_window(self.mask, rows, cols).reshape(self.N_samples, topk_min, 1, 2 * hh, 2 * wh)
        else:
            ps = self._patch_size
            select_inds = _choice(self._unfold_r0.shape[0], self.N_samples * topk, self.device)
            sel = torch.as_tensor(select_inds, device=self.device)
            r0s, c0s = self._unfold_r0[sel], self._unfold_c0[sel]
            select_img_patches = _crop(self._unfold_img, r0s, c0s, ps, ps).reshape(self.N_samples, topk, 3, ps, ps)
            select_mask_patches = _crop(self._unfold_mask, r0s, c0s, ps, ps).reshape(self.N_samples, topk, 1, ps, ps)
            weight_topks, topk_min = None, topk

        select_img_patches = select_img_patches.permute(0, 1, 3, 4, 2)
        select_mask_patches = select_mask_patches.permute(0, 1, 3, 4, 2)
        return select_img_patches, select_mask_patches, weight_topks, topk_min

    def sample_patch_fake(self, mode):
        pool = self.pool_train if mode == 'train' else self.pool_val
        hh, wh = self.patch_size_h_half, self.patch_size_w_half
        select_inds = _choice(pool.shape[0], self.N_samples, pool.device)
        pool_h = self._pool_host.get(mode)
        self._last_cent_host = None
        if pool_h is not None and self.img.dtype == torch.float32:
            # integer centroids: the glimpse rows / columns are c - size/2 + [0, size) (equal to the fp32 arithmetic of
            # _nearest_index for every integer c, tests/test_sampler_parity.py), built on the host and cropped by one
            # kernel per tensor
            cent_h = pool_h[np.asarray(select_inds)]
            rows = torch.as_tensor((cent_h[:, 0] - hh)[:, None] + np.arange(2 * hh)[None, :], device=self.device)
            cols = torch.as_tensor((cent_h[:, 1] - wh)[:, None] + np.arange(2 * wh)[None, :], device=self.device)
            select_patch_grids = torch.stack([rows[:, :, None].expand(-1, -1, 2 * wh),
                                              cols[:, None, :].expand(-1, 2 * hh, -1)], dim=-1)
            self._last_cent_host = cent_h
            return _gather_window(self.img, rows, cols), _gather_window(self.mask, rows, cols), select_patch_grids
        select_centroid = pool[torch.as_tensor(select_inds, device=pool.device)]
        cent = select_centroid.long()                       # int(left_h) of the reference truncates the same way
        r0, c0 = cent[:, 0] - hh, cent[:, 1] - wh
        rows = r0[:, None] + torch.arange(2 * hh, device=self.device)[None, :]
        cols = c0[:, None] + torch.arange(2 * wh, device=self.device)[None, :]
        select_patch_grids = torch.stack([rows[:, :, None].expand(-1, -1, 2 * wh),
                                          cols[:, None, :].expand(-1, 2 * hh, -1)], dim=-1)
        # glimpses are taken at the (possibly fractional) centroid like the reference does
        off = select_centroid.flip((1,))
        select_patch = extract_glimpse(self.img, (2 * hh, 2 * wh), off)
        select_patch_mask = extract_glimpse(self.mask, (2 * hh, 2 * wh), off)
        return select_patch, select_patch_mask, select_patch_grids

    def sample_patches(self, topk, invalid_ratio):
        prob = _uniform(self.device)
        if prob < 0.5:
            patch_source = 'val'
            fake, fake_mask, coords = self.sample_patch_fake('val')
            real, real_mask, weight_topk, topk = self.sample_patch_real(coords, topk=topk, invalid_ratio=invalid_ratio)
        elif prob > 0.5 and prob < 0.8:
            patch_source = 'train'
            fake, fake_mask, coords = self.sample_patch_fake('train')
            real, real_mask, weight_topk, topk = self.sample_patch_real(coords, topk=topk, invalid_ratio=invalid_ratio)
        else:
            fake, fake_mask, coords = self.sample_patch_fake('train')
            real, real_mask = fake.clone().permute(0, 2, 3, 1)[:, None], fake_mask.clone().permute(0, 2, 3, 1)[:, None]
            patch_source = 'same'
            topk = 1
            weight_topk = torch.ones(self.N_samples, dtype=torch.float32, device=self.device)
        if topk == 0:
            return None, None, None, None, None, None, topk, None
        if fake_mask is not None:
            fake = fake[:, None].tile([1, topk, 1, 1, 1])
            fake_mask = fake_mask[:, None].tile([1, topk, 1, 1, 1])
        return real, real_mask, fake, fake_mask, coords, patch_source, topk, weight_topk
