"""Host-side sequencing of the backbone kernels (eval-mode folded-BN pass and the train-mode batch-stat pass).

The arithmetic is in libsrb200.so (sr_pack_input / sr_pack_weight / sr_bn_fold / sr_conv / sr_bn_finalize /
sr_bn_apply); this module only owns buffers and launch order.  Reference semantics: BasicBlock.forward
(models/resnet_language.py:268-301), ResNet.forward (:170-181), nn.BatchNorm2d eval/train.
"""
import os
import threading
import time

import torch
import torch.nn.functional as F

from . import _lib as L
from . import device_rng
from . import host_rng
from . import rng
from . import ops

SLOPE = 0.1
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


_PINNED_POOL = {}
_MASK_POOL = {}     # device copies of the keep-masks: (run thread, 'mask', block, forward parity) -> grow-only uint8 buffer
_SCRATCH_POOL = {}
_SIDE_POOL = {}     # (run thread, device index) -> [mask side stream, sr_device_bernoulli workspace]: one per run thread for
                    # the life of the process - every new stream brings its own allocator pool, i.e. cudaMalloc calls (and
                    # the host stalls that come with them) in whatever sweep first uses it


def _pad16(c):
    return (c + 15) // 16 * 16


class BackboneEngine(object):
    """Packed (bf16, BN-folded) weights of one ResNet plus the two forward passes.

    `blocks` is the module's list of block descriptors: dict(prefix, mod, cin, cout, pool, downsample, drop_block,
    block_size) where `mod` carries conv1..3 / bn1..3 / downsample as parameter containers."""

    def __init__(self, blocks, chunk=1024, precision=None):
        self.blocks = blocks
        self.chunk = int(os.environ.get("SRB_EVAL_CHUNK", chunk))   # images per eval-mode pass (A/B timing override)
        # 'bf16': one tcgen05 pass on bf16 operands (throughput tier, features ~3e-3 from fp32);
        # 'bf16x3': error-compensated operand pairs, three passes into the same fp32 accumulator (parity tier, ~1e-5)
        self.precision = precision or os.environ.get("SRB_CONV_PRECISION", "bf16")
        if self.precision not in ("bf16", "bf16x3"):
            raise ValueError("srb200: conv precision must be 'bf16' or 'bf16x3', got %r" % (self.precision,))
        self._folded = None      # per block: dict(w1, s1, w2, s2, w3, s3[, wd])
        self._raw = None         # per block: unscaled packed weights (train-mode pass)
        self._fold_key = None
        self._staging = {}
        self._prefetch = None     # host_rng.MaskPrefetch for upcoming train-mode forwards
        self._db_ready = None     # DropBlock masks drawn ahead by start_dropblock_ahead: key -> (before, after, entry, kept, ...)
        self._db_done = True
        self._db_cv = threading.Condition()
        self._pf_fwd = 0          # index of the next train-mode forward relative to the prefetch plan
        self._pool_owner = threading.get_ident()   # staging buffers are pooled per run thread (helper threads use the owner's)
        self._mask_read = {}      # device mask buffer -> event after its last reader on the run's stream
        self._mask_pending = None
        self._dev_masks = None    # this forward's keep-masks when they are drawn on the device: block -> (keep, scale tensor | float)
        self._dev_plan = None     # (batch, input size, block counters, device) of a forward whose masks are not drawn yet
        self._dev_ready = None
        self._dev_fwd_done = {}   # forward parity -> event after that forward's last mask reader (buffer reuse fence)

    # ---------------------------------------------------------------- weight packing
    def _bn_key(self):
        key = []
        for b in self.blocks:
            m = b['mod']
            bns = [m.bn1, m.bn2, m.bn3] + ([m.downsample[1]] if b['downsample'] else [])
            convs = [m.conv1, m.conv2, m.conv3] + ([m.downsample[0]] if b['downsample'] else [])
            for bn in bns:
                key += [bn.running_mean._version, bn.running_var._version, bn.weight._version, bn.bias._version,
                        bn.running_mean.data_ptr()]
            for cv in convs:
                key += [cv.weight._version, cv.weight.data_ptr()]
        return tuple(key)

    def invalidate(self):
        self._fold_key = None

    @property
    def split(self):
        return self.precision == "bf16x3"

    def set_precision(self, precision):
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("srb200: conv precision must be 'bf16' or 'bf16x3', got %r" % (precision,))
        if precision != self.precision:
            self.precision = precision
            self._folded = self._raw = self._fold_key = None

    def _ensure_folded(self):
        key = self._bn_key()
        if self._folded is not None and key == self._fold_key:
            return
        folded = []
        for b in self.blocks:
            m = b['mod']
            d = {}
            for i, (cv, bn) in enumerate(((m.conv1, m.bn1), (m.conv2, m.bn2), (m.conv3, m.bn3))):
                scale, shift = ops.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, BN_EPS)
                d['w%d' % (i + 1)] = ops.pack_weight(cv.weight.detach(), scale, _pad16(cv.weight.shape[1]), split=self.split)
                d['s%d' % (i + 1)] = shift
            if b['downsample']:
                cv, bn = m.downsample[0], m.downsample[1]
                scale, shift = ops.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, BN_EPS)
                d['wd'] = ops.pack_weight(cv.weight.detach(), scale, _pad16(cv.weight.shape[1]), split=self.split)
                d['s3'] = d['s3'] + shift      # both branches land in one accumulator: shifts add
            folded.append(d)
        self._folded = folded
        self._fold_key = key

    def _ensure_raw(self):
        key = tuple(cv.weight._version for b in self.blocks for cv in
                    [b['mod'].conv1, b['mod'].conv2, b['mod'].conv3] + ([b['mod'].downsample[0]] if b['downsample'] else []))
        if self._raw is not None and self._raw[0] == key:
            return
        raw = []
        for b in self.blocks:
            m = b['mod']
            d = {('w%d' % (i + 1)): ops.pack_weight(cv.weight.detach(), None, _pad16(cv.weight.shape[1]), split=self.split)
                 for i, cv in enumerate((m.conv1, m.conv2, m.conv3))}
            if b['downsample']:
                d['wd'] = ops.pack_weight(m.downsample[0].weight.detach(), None, _pad16(m.downsample[0].weight.shape[1]),
                                          split=self.split)
            raw.append(d)
        self._raw = (key, raw)

    # ---------------------------------------------------------------- eval-mode pass
    def eval_features(self, x, taps=None):
        """x: CUDA fp32 NCHW [B,3,84,84] -> fp32 [B,640].  `taps`: optional list that receives the four stage
        outputs (NHWC bf16) for ResNet.forward(is_feat=True)."""
        self._ensure_folded()
        outs = []
        n = x.shape[0]
        n_chunks = (n + self.chunk - 1) // self.chunk
        step = (n + n_chunks - 1) // n_chunks          # equal chunks: no small last chunk with a ragged final wave
        for i0 in range(0, n, step):
            outs.append(self._eval_chunk(x[i0:i0 + step].contiguous(), taps))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def pack(self, x):
        """fp32 NCHW (already normalised, the reference's loader output) or uint8 NHWC (the raw image store: ToTensor +
        Normalize are fused into the packing kernel) -> NHWC bf16, 16 channels."""
        if x.dtype == torch.uint8:
            from dataset import transform_cfg
            return ops.pack_input_u8(x.contiguous(), transform_cfg.mean, transform_cfg.std, 16, split=self.split)
        return ops.pack_input(x.contiguous(), 16, split=self.split)

    def _eval_chunk(self, x, taps):
        h = self.pack(x)
        ev = getattr(self, 'conv_events', None)
        if ev is None:
            return self.eval_packed(h, taps)
        # measurement hook (bench.py): CUDA events around the convolution launches only, on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = self.eval_packed(h, taps)
        e1.record()
        ev.append((e0, e1, int(x.shape[0])))
        return out

    def eval_packed(self, h, taps=None):
        """The convolution launches of one eval-mode pass on an already packed NHWC bf16 input (18 for resnet18):
        what bench.py's roofline probe times."""
        self._ensure_folded()
        if taps is None:
            # one library call sequences the whole pass (18 launches for resnet18) out of a pooled workspace
            return ops.backbone_eval(h, self.blocks, self._folded, SLOPE)
        nb = len(self.blocks)
        for bi, (b, w) in enumerate(zip(self.blocks, self._folded)):
            cout = b['cout']
            last = bi == nb - 1
            h1 = ops.conv([(h, w['w1'])], cout, shift=w['s1'], slope=SLOPE, epilogue=L.SR_EPI_ACT)
            h2 = ops.conv([(h1, w['w2'])], cout, shift=w['s2'], slope=SLOPE, epilogue=L.SR_EPI_ACT)
            if b['downsample']:
                epi = L.SR_EPI_ACT_POOL2 if b['pool'] == 2 else L.SR_EPI_ACT
                out = ops.conv([(h2, w['w3']), (h, w['wd'])], cout, shift=w['s3'], slope=SLOPE, epilogue=epi)
                if last:                                    # resnet12: pooled last block, then AdaptiveAvgPool2d(1)
                    if taps is not None:
                        taps.append(out)
                    out = ops.global_avg(out)
            else:
                # the last block normally averages inside the conv epilogue; is_feat=True needs the map itself (f3)
                epi = L.SR_EPI_ACT_AVG if (last and taps is None) else L.SR_EPI_ACT
                out = ops.conv([(h2, w['w3'])], cout, shift=w['s3'], residual=h, slope=SLOPE, epilogue=epi)
                if last and taps is not None:
                    taps.append(out)
                    out = ops.global_avg(out)
            if taps is not None and not last and self.blocks[bi + 1]['prefix'].endswith('.0'):
                taps.append(out)                            # f0, f1, f2: outputs of layer1..3
            h = out
        return h

    # ---------------------------------------------------------------- train-mode pass (epoch 1 of a session)
    def device_masks(self, device=None):
        """True if this engine draws its keep-masks on the GPU (srb200.device_rng; SRB_MASKS=host or a failed self-check
        select the host generator threads below instead)."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return device_rng.available(device)

    def _draw_masks_on_device(self, B, size, counters, device):
        """All keep-masks of one train-mode forward, drawn on the device in forward order from the CPU generator's current
        state (which is moved past them): block index -> (keep uint8 NCHW, scale).  Runs on a side stream, so the ~0.3 ms of
        generator kernels overlap the first block's convolutions; returns None if a region layout is not supported."""
        plan = []
        for bi, b in enumerate(self.blocks):
            size = size // b['pool']
            dr = b.get('drop_rate', 0.1)
            if not dr > 0:
                continue
            shape = (B, b['cout'], size, size)
            if b['drop_block']:
                bs = b['block_size']
                gamma = self._dropblock_gamma(counters[b['prefix']], bs, size, dr)
                plan.append((bi, 1, gamma, (B, b['cout'], size - (bs - 1), size - (bs - 1)), shape, bs))
            else:
                plan.append((bi, 0, 1 - dr, shape, shape, 0))
        for k, (bi, kind, p, dshape, shape, bs) in enumerate(plan[:-1]):
            n = 1
            for d_ in dshape:
                n *= d_
            if device_rng.region_words(kind, n) & 1:
                return None
        main = torch.cuda.current_stream()
        side = self._side(device)
        cs = side[0]
        par = self._cur_fwd & 1
        out = {}
        with torch.cuda.stream(cs):
            prev = self._dev_fwd_done.get(par)
            if prev is not None:
                cs.wait_event(prev)            # the readers of these buffers two forwards ago
            regions = []
            post = []
            for bi, kind, p, dshape, shape, bs in plan:
                keep = self._dev_buf(('keep', bi, par), shape, device)
                if kind == 0:
                    regions.append((0, p, keep))
                    out[bi] = (keep, float(torch.ones(1).div_(p)))      # noise.div_(1 - drop_rate)
                else:
                    seeds = self._dev_buf(('seeds', bi, par), dshape, device)
                    scale = self._dev_buf(('scale', bi, par), (16,), device).view(torch.float32)
                    regions.append((1, p, seeds))
                    post.append((seeds, bs, keep, scale))
                    out[bi] = (keep, scale)
            side[1] = device_rng.draw(regions, side[1])
            for seeds, bs, keep, scale in post:
                device_rng.dropblock_keep(seeds, bs, keep, scale)
            ev = torch.cuda.Event()
            ev.record(cs)
        ops.LAUNCHES.add(3 + 3 * len(post))
        self._dev_ready = ev
        return out

    def _side(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (self._pool_owner, idx)
        ent = _SIDE_POOL.get(key)
        if ent is None:
            ent = _SIDE_POOL[key] = [torch.cuda.Stream(device=device), None]
        return ent

    def _dev_buf(self, key, shape, device):
        """Grow-only device uint8 buffer per (run thread, key), allocated on the mask stream's pool."""
        n = 1
        for d_ in shape:
            n *= int(d_)
        pkey = (self._pool_owner, 'dev') + tuple(key)
        buf = _MASK_POOL.get(pkey)
        if buf is None or buf.numel() < n:
            # (the old block returns to the mask stream's pool; that stream has already waited for its last readers)
            buf = torch.empty(1 << max(int(n - 1).bit_length(), 12), dtype=torch.uint8, device=device)
            _MASK_POOL[pkey] = buf
        return buf[:n].view(shape)

    def start_mask_prefetch(self, forwards):
        """Begin drawing, on a host thread, the dropout masks of the next len(forwards) train-mode forwards.
        forwards: [(skip_words_before, batch), ...] in the order they will run; `skip_words_before` generator draws are
        assumed to happen before that forward (the session's nn.Linear init).  DropBlock draws are skipped over (their
        gamma is not known yet) and made later from the live generator.  May cover every remaining session of a run."""
        if self.device_masks():
            return
        steps = []
        for f, (skip, batch) in enumerate(forwards):
            if skip:
                steps.append(('skip', skip))
            size = 84
            for bi, b in enumerate(self.blocks):
                size = size // b['pool']
                if not b.get('drop_rate', 0.1) > 0:      # BasicBlock.forward: `if self.drop_rate > 0` (resnet_language.py:292)
                    continue
                if b['drop_block']:
                    bs = b['block_size']
                    steps.append(('hold', (f, bi), batch * b['cout'] * (size - (bs - 1)) ** 2))
                else:
                    shape = (batch, b['cout'], size, size)
                    ent = self._pinned(('pf', f, bi), shape)      # waits for the last H2D out of this buffer
                    steps.append(('draw', (f, bi), ent[0], 1 - b.get('drop_rate', 0.1)))
        self._prefetch = host_rng.MaskPrefetch(steps)
        self._pf_fwd = 0

    @staticmethod
    def _dropblock_gamma(nbt, bs, size, drop_rate=0.1):
        keep_rate = max(1.0 - drop_rate / (20 * 2000) * nbt, 1.0 - drop_rate)
        return (1 - keep_rate) / bs ** 2 * size ** 2 / (size - bs + 1) ** 2

    def start_dropblock_ahead(self, forwards):
        """DropBlock keep-masks of the upcoming train-mode forwards on a second host thread.  forwards: [(batch,
        num_batches_tracked at that forward), ...] for the forwards the prefetch plan numbers self._pf_fwd, +1, ...  Their
        gamma depends on how many epochs the previous session ran, so they cannot be drawn with the dropout masks; but the
        prefetch thread keeps the generator state in front of every DropBlock region ('hold'), and from that state the
        seeds can be drawn off the live generator as soon as gamma is known - while the main thread is busy launching the
        first blocks.  Consumed (and verified against the live generator) in draw_mask."""
        if self.device_masks():
            return
        t = getattr(self, '_db_thread', None)
        if t is not None:
            t.join()
        self._db_ready = {}
        self._db_done = False
        if self._prefetch is None or not self._prefetch.alive():
            self._db_done = True
            return
        plan = [(self._pf_fwd + k, batch, nbt) for k, (batch, nbt) in enumerate(forwards)]
        pf = self._prefetch

        def work():
            try:
                for f, batch, nbt in plan:
                    size = 84
                    for bi, b in enumerate(self.blocks):
                        size = size // b['pool']
                        if not b['drop_block'] or not b.get('drop_rate', 0.1) > 0:
                            continue
                        st = pf.hold_state((f, bi))
                        if st is None:
                            return
                        bs = b['block_size']
                        gamma = self._dropblock_gamma(nbt, bs, size, b.get('drop_rate', 0.1))
                        shape = (batch, b['cout'], size, size)
                        seed_shape = (batch, b['cout'], size - (bs - 1), size - (bs - 1))
                        seeds = self._scratch(('seed_ahead', f & 1, bi), seed_shape)
                        state = st[0].clone()
                        if host_rng.replay_into(state, gamma, 1, seeds) < 0 or not torch.equal(state, st[1]):
                            return
                        ent = self._pinned(('dbm', f & 1, bi), shape)
                        kept = host_rng.dropblock_keep(seeds, bs, ent[0])
                        with self._db_cv:
                            self._db_ready[(f, bi)] = (st[0], st[1], ent, kept, gamma, shape)
                            self._db_cv.notify_all()
            finally:
                with self._db_cv:
                    self._db_done = True
                    self._db_cv.notify_all()

        self._db_thread = threading.Thread(target=work, daemon=True)
        self._db_thread.start()

    def _dropblock_take(self, key, gamma, shape):
        """-> (pinned entry, kept) drawn ahead for `key`, after moving the live generator past it; or None."""
        if getattr(self, '_db_ready', None) is None:
            return None
        t0 = time.perf_counter()
        with self._db_cv:
            while key not in self._db_ready and not self._db_done:
                self._db_cv.wait()
            got = self._db_ready.pop(key, None)
        host_rng.WAIT_S[0] += time.perf_counter() - t0
        if got is None:
            return None
        before, after, ent, kept, g, shp = got
        if g != gamma or tuple(shp) != tuple(shape) or not torch.equal(rng.get_state(), before):
            return None
        rng.set_state(after)
        return ent, kept

    def mask_prefetch_alive(self):
        return self.device_masks() or (self._prefetch is not None and self._prefetch.alive())

    def _scratch(self, key, shape):
        """Reusable pageable uint8 host buffer (process-wide pool, power-of-two capacity)."""
        n = 1
        for d in shape:
            n *= int(d)
        key = (self._pool_owner,) + tuple(key)     # one pool per run thread: concurrent runs never share a buffer
        flat = _SCRATCH_POOL.get(key)
        if flat is None or flat.numel() < n:
            flat = torch.empty(1 << max(int(n - 1).bit_length(), 12), dtype=torch.uint8)
            _SCRATCH_POOL[key] = flat
        return flat[:n].view(shape)

    def _pinned(self, key, shape, sync=True):
        """Reusable pinned staging buffer (uint8) + the event of its last H2D copy."""
        n = 1
        for d in shape:
            n *= int(d)
        key = (self._pool_owner,) + tuple(key)     # one pool per run thread: concurrent runs never share a buffer
        ent = _PINNED_POOL.get(key)     # process-wide: cudaHostAlloc is far too slow to repeat per model instance
        if ent is None or ent[2].numel() < n:
            cap = 1 << max(int(n - 1).bit_length(), 12)          # power-of-two capacity: batches grow session by session
            flat = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            ent = [None, None, flat]
            _PINNED_POOL[key] = ent
        if sync and ent[1] is not None:
            ent[1].synchronize()     # the previous copy out of this buffer must have finished before it is rewritten
        ent[0] = ent[2][:n].view(shape)
        return ent

    def draw_mask(self, bi, batch, size, counters, device):
        """Keep-mask of block `bi` for one train-mode forward, drawn from torch's CPU generator exactly like the
        reference does at this point of the forward (F.dropout after layerX.0, DropBlock after layer3.1 / layer4.1;
        resnet_language.py:292-299, 311-325).  -> (keep uint8 NCHW on `device`, scale)."""
        b = self.blocks[bi]
        shape = (batch, b['cout'], size, size)
        drop_rate = b.get('drop_rate', 0.1)
        if not drop_rate > 0:                                                  # `if self.drop_rate > 0` (:292): no draw at all
            return None, 1.0
        if self._dev_plan is not None:
            # All masks of this forward are drawn on the device NOW, i.e. after the first block's convolutions have been
            # queued: the generator kernels (side stream) and the host-side jump of the CPU generator (~0.3 ms) then
            # overlap that block's device work instead of preceding it.
            self._dev_masks = self._draw_masks_on_device(*self._dev_plan)
            self._dev_plan = None
        if self._dev_masks is not None:
            if self._dev_ready is not None:
                torch.cuda.current_stream().wait_event(self._dev_ready)
                self._dev_ready = None
            return self._dev_masks[bi]
        if not b['drop_block']:
            scale = float(torch.ones(1).div_(1 - drop_rate))                   # noise.div_(1 - p)
            got = self._prefetch.take((self._cur_fwd, bi), shape) if self._prefetch is not None else None
            if got is not None:                                                # drawn ahead of time on the host thread
                ent = self._pinned(('pf', self._cur_fwd, bi), shape, sync=False)
            else:
                ent = self._pinned(('m', bi, self._cur_fwd & 1), shape)
                host_rng.bernoulli_u8(shape, 1 - drop_rate, 0, out=ent[0])    # noise.bernoulli_(1 - p)
        else:
            bs = b['block_size']
            nbt = counters[b['prefix']]
            gamma = self._dropblock_gamma(nbt, bs, size, drop_rate)
            ahead = self._dropblock_take((self._cur_fwd, bi), gamma, shape)
            if ahead is not None:                                              # drawn on the second host thread
                ent, kept = ahead
            else:
                seed_shape = (batch, b['cout'], size - (bs - 1), size - (bs - 1))
                seed_buf = self._scratch(('seed', bi), seed_shape)             # reused: a fresh tensor page-faults every time
                seeds, n_seed = host_rng.bernoulli_u8(seed_shape, gamma, 1, out=seed_buf)   # Bernoulli(gamma).sample(...)
                # alternate staging buffers: the second forward of a session must not wait for the H2D copy of the first
                ent = self._pinned(('m', bi, self._cur_fwd & 1), shape)
                kept = host_rng.dropblock_keep(seeds, bs, ent[0])              # _compute_block_mask, 1 - mask, .sum()
            # countM / count_ones: python int over an fp32 0-d tensor -> fp32 division
            scale = float(torch.tensor(float(ent[0].numel()), dtype=torch.float32) / torch.tensor(float(kept), dtype=torch.float32))
        # Upload on a side stream: the 20-40 MB mask copy then overlaps the convolutions of this block that are already
        # queued on the run's stream instead of sitting between them (43 MB per support forward, ~26 ms of copies per sweep).
        main = torch.cuda.current_stream()
        cs = self._side(device)[0]
        # The device copy lives in a grow-only buffer per (block, forward parity) instead of a fresh allocation: mask sizes
        # change with every forward (support / memory batches), so fresh allocations on the copy stream kept reaching
        # cudaMalloc, where the launching thread was seen to block for 100-300 ms.  The buffer's previous reader (the
        # sr_bn_apply of two forwards ago, on the run's stream) is fenced with an event before the copy overwrites it.
        dkey = ('mask', bi, self._cur_fwd & 1)
        pkey = (self._pool_owner,) + dkey
        n_el = ent[0].numel()
        buf = _MASK_POOL.get(pkey)
        last_read = self._mask_read.get(dkey)
        with torch.cuda.stream(cs):
            if buf is None or buf.numel() < n_el:
                # (allocated with the COPY stream current: the block then comes from that stream's pool, i.e. nothing still
                # queued on the run's stream can be using its memory when the copy below overwrites it)
                if buf is not None:
                    buf.record_stream(main)        # its last reader runs on the run's stream
                buf = torch.empty(1 << max(int(n_el - 1).bit_length(), 12), dtype=torch.uint8, device=device)
                _MASK_POOL[pkey] = buf
            keep = buf[:n_el].view(shape)
            if last_read is not None:
                cs.wait_event(last_read)
            keep.copy_(ent[0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        ent[1] = ev                    # the staging buffer may be rewritten once this copy has finished
        main.wait_event(ev)            # consumers on the run's stream (sr_bn_apply) wait for the copy only
        self._mask_pending = dkey      # train_features records the reader's event after the consuming launch
        return keep, scale

    def train_features(self, x, counters):
        """One train-mode forward (batch-stat BN, running-stat EMA in place, dropout / DropBlock) -> fp32 [B,640]."""
        self._ensure_raw()
        raw_w = self._raw[1]
        B = x.shape[0]
        dev = x.device
        size = x.shape[2] if x.dtype != torch.uint8 else x.shape[1]
        self._cur_fwd = self._pf_fwd
        self._pf_fwd += 1
        self._dev_masks = None
        self._dev_plan = (B, size, counters, dev) if self.device_masks(dev) else None
        h = self.pack(x)
        nb = len(self.blocks)

        # one zeroed fp64 scratch for the (sum, sum of squares) of every conv of this forward, one fused counter bump at
        # the end: ~40 fewer tiny launches per forward (the train-mode pass is host-bound)
        n_stats = sum(2 * b['cout'] * (4 if b['downsample'] else 3) for b in self.blocks)
        stats_all = torch.zeros(n_stats, dtype=torch.float64, device=dev)
        stats_off = [0]
        bumped = []

        for bi, (b, w) in enumerate(zip(self.blocks, raw_w)):
            m, cout = b['mod'], b['cout']
            last = bi == nb - 1
            ds = bool(b['downsample'])
            bns = [m.bn1, m.bn2, m.bn3] + ([m.downsample[1]] if ds else [])
            n_st = 2 * cout * len(bns)
            stats = stats_all[stats_off[0]:stats_off[0] + n_st]
            stats_off[0] += n_st
            # conv1 .. conv3 (+ downsample conv), their batch statistics / running-stat updates and the two inner BN +
            # LeakyReLU passes: one library call (sr_train_block) instead of ~11 launches sequenced from Python
            r3, rd, mi = ops.train_block(
                h, [w['w1'], w['w2'], w['w3']] + ([w['wd']] if ds else []),
                [(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var) for bn in bns], stats,
                h.shape[-1], cout, ds, BN_EPS, BN_MOMENTUM, SLOPE)
            bumped += [bn.num_batches_tracked for bn in bns]
            size = size // b['pool']
            # drawn here, in forward order: the host RNG of block i overlaps the GPU work already queued for block i
            keep, scale = self.draw_mask(bi, B, size, counters, dev)
            pool = -1 if last and b['pool'] == 1 else (2 if b['pool'] == 2 else 0)
            if ds:
                bnd = m.downsample[1]
                h = ops.bn_apply(r3, mi[2, 0], mi[2, 1], m.bn3.weight.detach(), m.bn3.bias.detach(), res_raw=rd,
                                 res_bn=(mi[3, 0], mi[3, 1], bnd.weight.detach(), bnd.bias.detach()), lrelu=True, slope=SLOPE,
                                 pool=pool, keep=keep, keep_scale=scale, split=self.split)
            else:
                h = ops.bn_apply(r3, mi[2, 0], mi[2, 1], m.bn3.weight.detach(), m.bn3.bias.detach(), res_act=h, lrelu=True,
                                 slope=SLOPE, pool=pool, keep=keep, keep_scale=scale, split=self.split)
            if self._mask_pending is not None:      # the mask buffer of this block has been read: fence its next overwrite
                ev = torch.cuda.Event()
                ev.record()
                self._mask_read[self._mask_pending] = ev
                self._mask_pending = None
        if self._dev_masks is not None:             # fence for the mask buffers of this parity (rewritten two forwards on)
            ev = torch.cuda.Event()
            ev.record()
            self._dev_fwd_done[self._cur_fwd & 1] = ev
            self._dev_masks = None
        torch._foreach_add_(bumped, 1)   # BatchNorm2d.num_batches_tracked of every BN that ran
        self.invalidate()   # running statistics moved: the folded weights are stale
        if h.dim() >= 4:    # resnet12: the last block is pooled 2x2, the global average follows
            h = ops.global_avg(h)
        return h
