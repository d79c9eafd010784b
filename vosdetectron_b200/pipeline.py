"""Frame-batched, device-resident region pipeline and the frame-sharded clip driver.

The reference runs the path once per frame with three host round trips
(lib/core/test.py:50-120 -> model_builder.py:146-250).  Here a batch of frames goes through

    generate_proposals (all levels x frames, 3 launches)  -> collect+distribute (1 launch)
    -> box RoIAlign over all levels (1 launch)  ...box head (not part of this library)...
    -> distribute(detections) (1 launch) -> mask RoIAlign (1 launch) ...mask head...
    -> paste (1 launch)

without leaving the GPU.  Frames are independent units, so a clip is
split frame-wise over ranks exactly like the reference's multi-GPU test splits its dataset
(np.array_split, lib/utils/subprocess.py:56), and the only collective is the final all-gather
of per-frame detections and (bit-packed) masks.
"""
import numpy as np
import torch

from . import ops
from .config import RegionConfig
from .modeling.generate_anchors import fpn_level_anchors


class RegionPipeline:
    def __init__(self, cfg=None, training=False, box_resolution=7, mask_resolution=14, sampling_ratio=2):
        self.cfg = cfg or RegionConfig()
        self.training = training
        self.box_resolution = box_resolution
        self.mask_resolution = mask_resolution
        self.sampling_ratio = sampling_ratio
        c = self.cfg
        self.rpn_levels = list(range(c.rpn_min_level, c.rpn_max_level + 1))
        self.roi_levels = list(range(c.roi_min_level, c.roi_max_level + 1))
        self.anchors = {l: fpn_level_anchors(l, c.rpn_anchor_start_size, c.rpn_aspect_ratios, c.rpn_min_level)
                        for l in self.rpn_levels}
        self._ws = None
        self.overlap = True
        # payload of a frame-sharded clip's all-gather, produced inside step(): False = none, True = 1 bit per pixel
        # (written by the paste kernel itself), "rle" = COCO RLE strings (fused paste -> RLE kernel: what the reference
        # ships between its per-GPU processes, ~1 KB per detection instead of 51 KB)
        self.packed_masks = False
        self._rle = None
        self.rle_bytes_per_det = 4096  # capacity of the gathered string arena (average per detection)
        self._side = None

    # ---- stage 1: proposals -> top RoIs + FPN levels (rows of frame f are in group f) ----------
    def proposals(self, rpn, im_info, images_per_group=1, mark=None):
        """rpn: {lvl: (scores (B,A,H,W), deltas (B,4A,H,W))} CUDA tensors; im_info (B,3) CUDA."""
        m = self.cfg.mode(self.training)
        inputs = [(rpn[l][0], rpn[l][1], self.anchors[l], float(2 ** l)) for l in self.rpn_levels]
        rois, probs, count = ops.generate_proposals_cuda(inputs, im_info, m.pre_nms_topN, m.post_nms_topN,
                                                         m.nms_thresh, m.min_size, zero_fill=False)   # collect honours count
        if mark:
            mark("collect_distribute")
        c = self.cfg
        return ops.collect_distribute_cuda(rois, probs, count, c.collect_post_topN(self.training),
                                           images_per_group, c.roi_min_level, c.roi_max_level,
                                           c.roi_canonical_scale, c.roi_canonical_level)

    # ---- stage 2: multi-level RoIAlign on device-resident rois --------------------------------
    def roi_features(self, feats, rois, level, resolution):
        """feats: {lvl: (B,C,H,W)}; rois (R,5) with column 0 = frame index in the batch;
        level (R) int32 FPN level in [roi_min_level, roi_max_level]."""
        k_min = self.cfg.roi_min_level
        fl = [feats[l] for l in self.roi_levels]
        sc = [1.0 / 2 ** l for l in self.roi_levels]
        return ops.roi_align_ml_forward(fl, sc, rois, (level - k_min).to(torch.int32), resolution, resolution,
                                        self.sampling_ratio)

    # ---- whole step on device-resident inputs -------------------------------------------------
    def _mask_branch(self, feats, det_boxes, det_cls, det_masks, frame_hw, im_scale, mark):
        """detections -> blob coords -> level -> mask RoIAlign (im_detect_mask, test.py:366-402), and the
        paste of the mask-head output (segm_results, test.py:801-855)."""
        c = self.cfg
        B, D = det_boxes.shape[:2]
        rle = None
        if self.packed_masks == "rle":
            # First kernel of this chain: the fused paste -> RLE kernel is latency-bound (one CTA per detection, the
            # largest boxes are its tail), like the proposal chain that starts on the other stream at the same time,
            # so the two share the GPU before the bandwidth-bound kernels arrive.
            mark("paste_rle")
            K0, M0 = det_masks.shape[2], det_masks.shape[3]
            cap = B * D * self.rle_bytes_per_det
            rle = ops.paste_rle_cuda(det_masks.view(B * D, K0, M0, M0),
                                     det_cls.view(-1) if (c.mrcnn_cls_specific_mask and K0 > 1) else None, det_boxes.view(B * D, 4),
                                     frame_hw[0], frame_hw[1], c.mrcnn_thresh_binarize, run_capacity=cap // 2,
                                     str_capacity=cap)
        mark("mask_rois")
        frame_idx = torch.arange(B, device=det_boxes.device, dtype=torch.float32).view(B, 1, 1).expand(B, D, 1)
        mask_rois = torch.cat([frame_idx, det_boxes * im_scale], dim=2).view(B * D, 5).contiguous()
        mlevel, _, _, _ = ops.distribute_cuda(mask_rois, c.roi_min_level, c.roi_max_level,
                                              c.roi_canonical_scale, c.roi_canonical_level)
        mark("roialign_mask")
        mask_feats = self.roi_features(feats, mask_rois, mlevel, self.mask_resolution)
        mark("paste")
        K, M = det_masks.shape[2], det_masks.shape[3]
        # (B,D,1,M,M): the caller already selected every detection's class channel (HostPipeline does, on the host,
        # so that 1 of the 81 channels crosses PCIe): paste reads channel 0, as for a class-agnostic mask head
        cls = det_cls.view(-1) if (c.mrcnn_cls_specific_mask and K > 1) else None
        if self.packed_masks == "rle":
            pasted = ops.paste_masks_cuda(det_masks.view(B * D, K, M, M), cls, det_boxes.view(B * D, 4),
                                          frame_hw[0], frame_hw[1], c.mrcnn_thresh_binarize)
            packed = None
        elif self.packed_masks:
            # the paste kernel also writes the 1-bit-per-pixel copy (the all-gather payload of a sharded clip)
            pasted, packed = ops.paste_masks_packed_cuda(det_masks.view(B * D, K, M, M), cls, det_boxes.view(B * D, 4),
                                                         frame_hw[0], frame_hw[1], c.mrcnn_thresh_binarize)
        else:
            pasted = ops.paste_masks_cuda(det_masks.view(B * D, K, M, M), cls, det_boxes.view(B * D, 4),
                                          frame_hw[0], frame_hw[1], c.mrcnn_thresh_binarize)
            packed = None
        mark("mask_end")
        self._rle = rle
        return mask_rois, mlevel, mask_feats, pasted, packed

    def step(self, rpn, im_info, feats, det_boxes, det_cls, det_masks, frame_hw, im_scale, mark=None, overlap=None):
        """One batch of B frames.
        det_boxes (B,D,4) original-frame coords, det_cls (B,D) int32, det_masks (B,D,K,M,M):
        the box-head / mask-head outputs the pipeline sits between (synthetic in the benchmark).
        ``mark(name)`` (optional) is called on the launching stream right before each stage
        (bench.py records CUDA events there).  Returns dict of device tensors.

        The step has two chains that only meet through the heads outside this library: proposals ->
        collect -> box RoIAlign, and mask RoIs -> mask RoIAlign / paste.  The proposal chain is a
        latency-bound sequence on a few SMs (one cluster per (level, frame) segment, one CTA per
        segment in the NMS reduce), so with ``overlap`` (default: ``self.overlap``) it runs on a second,
        high-priority stream beside the mask chain and both join before the box RoIAlign, which then has
        the GPU to itself."""
        c = self.cfg
        mark = mark or (lambda name: None)
        overlap = self.overlap if overlap is None else overlap
        B, D = det_boxes.shape[:2]
        main = torch.cuda.current_stream()

        def proposal_chain():
            mark("proposals")
            prop = self.proposals(rpn, im_info, mark=mark)
            # rows beyond count[g] are zero boxes on frame 0 level k_min: harmless filler, masked by count
            level = prop["level"].view(-1).clamp_(c.roi_min_level, c.roi_max_level)
            mark("proposals_end")
            return prop, level

        if overlap:
            # The proposal chain goes to a HIGH-PRIORITY second stream: its few CTAs (clusters of 4 per segment,
            # one CTA per segment in the NMS reduce) are placed ahead of the pending CTAs of the mask RoIAlign /
            # paste kernels that fill the rest of the GPU from the caller's stream.
            if self._side is None:
                self._side = torch.cuda.Stream(det_boxes.device, priority=-1)
            side = self._side
            side.wait_stream(main)
            with torch.cuda.stream(side):
                prop, level = proposal_chain()
                if overlap == "full":
                    # the box RoIAlign follows the proposals on the side stream, beside mask RoIAlign + paste
                    # (bench: 1.296 -> 1.212 ms per step; with overlap=True both chains join first and it runs alone)
                    post = prop["rois"].shape[1]
                    mark("roialign_box")
                    box_feats = self.roi_features(feats, prop["rois"].view(B * post, 5), level, self.box_resolution)
                    mark("end")
            # No record_stream on the side stream's tensors: it would park every freed block behind an event and
            # make the caching allocator cudaMalloc fresh outputs each step.  Reuse is safe without it: the blocks
            # are only re-allocated by the side stream, whose next use starts with wait_stream(main) above.
            mask_rois, mlevel, mask_feats, pasted, packed = self._mask_branch(feats, det_boxes, det_cls, det_masks,
                                                                      frame_hw, im_scale, mark)
            mark("join_wait")
            main.wait_stream(side)
        else:
            prop, level = proposal_chain()
        if overlap != "full":
            post = prop["rois"].shape[1]
            rois = prop["rois"].view(B * post, 5)
            mark("roialign_box")
            box_feats = self.roi_features(feats, rois, level, self.box_resolution)
            mark("end")
        if not overlap:
            mask_rois, mlevel, mask_feats, pasted, packed = self._mask_branch(feats, det_boxes, det_cls, det_masks,
                                                                      frame_hw, im_scale, mark)
        return {"rois": prop["rois"], "roi_count": prop["count"], "roi_level": prop["level"],
                "box_feats": box_feats, "mask_feats": mask_feats, "mask_rois": mask_rois, "mask_level": mlevel,
                "masks": pasted.view(B, D, frame_hw[0], frame_hw[1]),
                "masks_packed": None if packed is None else packed.view(B, D, -1),
                # "rle" payload: chars (B, cap/B) uint8 arena slices are NOT per frame -- offsets index the whole arena
                "masks_rle": self._rle}


    def run_clip(self, batches, frame_hw, im_scale, group=None, transport="auto"):
        """This rank's share of a frame-sharded clip: ``batches`` = dicts of device tensors (keys of
        HostPipeline.INPUT_KEYS); per-frame detection records [x1,y1,x2,y2,score,cls] and the 1-bit-per-pixel pasted
        masks of every batch are all-gathered behind the next batch's kernels (FrameGather).  Returns (dets, masks
        packed) of the whole clip on every rank, see ``run_clip``."""
        keep, self.packed_masks = self.packed_masks, True

        def step_fn(b):
            return self.step(b["rpn"], b["im_info"], b["feats"], b["det_boxes"], b["det_cls"], b["det_masks"], frame_hw, im_scale)

        def records_fn(b, out):
            cls = b["det_cls"].unsqueeze(-1).float()
            score = b["det_scores"].unsqueeze(-1) if "det_scores" in b else torch.ones_like(cls)
            return torch.cat([b["det_boxes"], score, cls], dim=2), out["masks_packed"]
        try:
            return run_clip(batches, step_fn, records_fn, group)
        finally:
            self.packed_masks = keep


# Number of library kernels one RegionPipeline.step enqueues (counted, see bench.py):
#   topk_decode + nms_mask + nms_reduce + collect_distribute + roialign(box) + distribute +
#   roialign(mask) + paste
STEP_LAUNCHES = 8


def shard_frames(num_frames, world_size, rank):
    """Frames of rank `rank`: contiguous np.array_split, the reference's convention
    (lib/utils/subprocess.py:56)."""
    return np.array_split(np.arange(num_frames), world_size)[rank]


def pack_mask_bits(masks_u8):
    """(…,H,W) uint8 {0,1} -> (…, ceil(H*W/8)) uint8, 8 pixels per byte (lossless; the payload
    of the all-gather)."""
    if masks_u8.is_cuda:
        return ops.pack_mask_bits_cuda(masks_u8)          # one streaming kernel
    flat = masks_u8.reshape(*masks_u8.shape[:-2], -1)     # CPU (gloo tests): same layout with torch ops
    n = flat.shape[-1]
    pad = (-n) % 8
    if pad:
        flat = torch.nn.functional.pad(flat, (0, pad))
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=flat.device)
    return (flat.view(*flat.shape[:-1], -1, 8) * w).sum(dim=-1, dtype=torch.uint8)


def unpack_mask_bits(packed, h, w):
    sh = torch.arange(8, dtype=torch.uint8, device=packed.device)
    bits = (packed.unsqueeze(-1) >> sh) & 1
    return bits.reshape(*packed.shape[:-1], -1)[..., :h * w].reshape(*packed.shape[:-1], h, w)


def all_gather_frames(dets, masks, group=None, async_op=False):
    """All-gather per-frame detection records (F_r, D, 6) fp32 and masks (F_r, D, …) uint8 over the
    ranks of `group` (NCCL on GPUs, gloo on CPU in the tests).  Every rank must hold the same
    number of frames (pad the clip); returns tensors with F = F_r * world frames in rank order.
    With ``async_op`` the collectives run on NCCL's stream behind the kernels already enqueued and a
    third value, the list of work handles, is returned: call ``.wait()`` on them before reading the
    outputs -- the next batch's kernels overlap the gather."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    d_out = torch.empty((world * dets.shape[0],) + tuple(dets.shape[1:]), dtype=dets.dtype, device=dets.device)
    m_out = torch.empty((world * masks.shape[0],) + tuple(masks.shape[1:]), dtype=masks.dtype, device=masks.device)
    w1 = dist.all_gather_into_tensor(d_out, dets.contiguous(), group=group, async_op=async_op)
    w2 = dist.all_gather_into_tensor(m_out, masks.contiguous(), group=group, async_op=async_op)
    if async_op:
        return d_out, m_out, [w1, w2]
    return d_out, m_out


class FrameGather:
    """All-gather of one batch's per-frame detection records and bit-packed masks over the ranks of a frame-sharded
    clip (the single exchange of the path; reference twin: the per-range result files merged by
    lib/core/test_engine.py:168-213).

    transport "ce" (default on GPUs when symmetric memory is available): every rank PUSHES its payload into each
    peer's receive buffer with copy-engine peer copies over NVLink (``torch.distributed._symmetric_memory``: the
    buffers of all ranks are mapped into every process), bracketed by two signal-pad barriers (one-CTA kernels).  No
    NCCL kernel holds SMs while the next batch's RoIAlign runs -- the limiter of the round-1 scaling (rank 0's box
    RoIAlign 0.77 -> 0.98 ms at 8 GPUs).  transport "nccl": ``all_gather_into_tensor`` (also the gloo path of the CPU
    tests).  Two rotating slots: the gather of batch i overlaps the kernels of batch i+1.

        g = FrameGather(dets.shape, dets.dtype, masks.shape, masks.dtype, device)
        h = g.start(dets, masks)       # async; returns the slot
        d_all, m_all = g.finish(h)     # (world * F_r, ...) tensors in rank order: views of the receive slot, valid
                                       # until start() reuses slot h (consume or copy them on the current stream before)
    """

    def __init__(self, dets_shape, dets_dtype, masks_shape, masks_dtype, device, group=None, transport="auto", slots=2):
        import torch.distributed as dist
        self.dist, self.group, self.device, self.slots = dist, group, torch.device(device), slots
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.d_shape, self.d_dtype = tuple(dets_shape), dets_dtype
        self.m_shape, self.m_dtype = tuple(masks_shape), masks_dtype
        self.d_bytes = int(np.prod(self.d_shape)) * torch.empty((), dtype=dets_dtype).element_size()
        self.m_bytes = int(np.prod(self.m_shape)) * torch.empty((), dtype=masks_dtype).element_size()
        self.pay = (self.d_bytes + self.m_bytes + 255) // 256 * 256        # one rank's payload, 256-byte aligned
        self.transport, self.why = "nccl", None
        self.count = 0
        self.works = [None] * slots
        self.outs = [None] * slots
        self.released = [True] * slots
        if transport in ("auto", "ce") and self.device.type == "cuda" and self.world > 1:
            try:
                self._init_ce()
                self.transport = "ce"
            except Exception as exc:  # noqa: BLE001  (no P2P / no fabric handles in this container: NCCL instead)
                self.why = "%s: %s" % (type(exc).__name__, str(exc)[:160])
                if transport == "ce":
                    raise

    # ---- copy-engine transport -----------------------------------------------------------------
    # Symmetric buffer of every rank: [slot][source rank][payload] + flag words.  All signalling uses copy engines and
    # stream memory operations, never a kernel: a spinning barrier kernel cannot be scheduled beside the persistent
    # RoIAlign kernel (one CTA per SM, all registers), which would serialise the exchange behind it.
    #   data flag  F[slot][src] (in the RECEIVER's buffer): src copies the sequence number there after its payload
    #   ack  flag  A[slot][dst] (in the SENDER's buffer):   dst copies the sequence number there once it has read the slot
    def _init_ce(self):
        import torch.distributed._symmetric_memory as symm_mem
        from cuda.bindings import driver as drv
        self.drv = drv
        dist = self.dist
        gname = (self.group or dist.group.WORLD).group_name
        try:
            symm_mem.enable_symm_mem_for_group(gname)
        except Exception:  # noqa: BLE001  (newer releases enable it implicitly)
            pass
        W, S = self.world, self.slots
        dpad, mpad = (self.d_bytes + 15) // 16 * 16, (self.m_bytes + 15) // 16 * 16     # usually no padding: results are views
        slot_bytes = W * (dpad + mpad)                           # [dets of every rank][masks of every rank]
        self.data_bytes = S * slot_bytes
        total = self.data_bytes + 2 * S * W * 4                  # + uint32 data flags, ack flags
        self.recv = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
        self.recv.zero_()
        self.hdl = symm_mem.rendezvous(self.recv, gname)
        peer = [self.hdl.get_buffer(r, (total,), torch.uint8) for r in range(W)]
        self.seq_vals = torch.arange(1 << 16, dtype=torch.int32, device=self.device).view(torch.uint8)   # constants to copy from
        self.cs = torch.cuda.Stream(self.device)
        self.ev_ready = [torch.cuda.Event() for _ in range(S)]
        self.ev_staged = [torch.cuda.Event() for _ in range(S)]
        self.ev_done = [torch.cuda.Event() for _ in range(S)]
        self.seq = [0] * S
        self.stage_d = [torch.empty(self.d_bytes, dtype=torch.uint8, device=self.device) for _ in range(S)]
        self.stage_m = [torch.empty(self.m_bytes, dtype=torch.uint8, device=self.device) for _ in range(S)]

        def flag(buf, kind, slot, r):
            o = self.data_bytes + 4 * ((kind * S + slot) * W + r)
            return buf[o:o + 4]
        # every view the per-batch path touches is built once (the per-batch host cost is what limits a 1 ms step)
        order = [(self.rank + 1 + k) % W for k in range(W)]      # staggered: rank r starts with peer r + 1
        self.push = []           # [slot] -> list of (dst dets view, dst masks view, dst data flag, local ack flag address)
        self.out_d, self.out_m, self.wait_addr, self.ack_dst = [], [], [], []
        base_ptr = self.recv.data_ptr()
        for sl in range(S):
            b0 = sl * slot_bytes
            self.push.append([(peer[r][b0 + self.rank * dpad:b0 + self.rank * dpad + self.d_bytes],
                               peer[r][b0 + W * dpad + self.rank * mpad:b0 + W * dpad + self.rank * mpad + self.m_bytes],
                               flag(peer[r], 0, sl, self.rank),
                               base_ptr + self.data_bytes + 4 * ((1 * S + sl) * W + r)) for r in order])
            if dpad == self.d_bytes:
                d = self.recv[b0:b0 + W * dpad]
            else:
                d = self.recv[b0:b0 + W * dpad].view(W, dpad)[:, :self.d_bytes]
            if mpad == self.m_bytes:
                m = self.recv[b0 + W * dpad:b0 + W * (dpad + mpad)]
            else:
                m = self.recv[b0 + W * dpad:b0 + W * (dpad + mpad)].view(W, mpad)[:, :self.m_bytes]
            self.out_d.append(d)
            self.out_m.append(m)
            self.wait_addr.append([base_ptr + self.data_bytes + 4 * ((0 * S + sl) * W + r) for r in range(W)])
            self.ack_dst.append([flag(peer[r], 1, sl, self.rank) for r in range(W)])
        torch.cuda.synchronize(self.device)
        self.hdl.barrier(channel=0)                              # everybody's flags are zero before the first push
        torch.cuda.synchronize(self.device)

    def _wait_flag(self, stream_handle, addr, value):
        err, = self.drv.cuStreamWaitValue32(stream_handle, addr, value, self.drv.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ)
        if int(err) != 0:
            raise RuntimeError("cuStreamWaitValue32 failed: %s" % err)

    def _start_ce(self, dets, masks, slot, stage=True):
        main = torch.cuda.current_stream(self.device)
        self.ev_ready[slot].record(main)
        self.seq[slot] += 1
        seq = self.seq[slot]
        if seq >= (1 << 16) - 1:
            raise RuntimeError("FrameGather: sequence space exhausted (65534 gathers per slot); create a new one")
        val = self.seq_vals[4 * seq:4 * seq + 4]
        cs_h = self.cs.cuda_stream
        with torch.cuda.stream(self.cs):
            self.cs.wait_event(self.ev_ready[slot])
            if stage:
                sd, sm = self.stage_d[slot], self.stage_m[slot]
                sd.copy_(dets.contiguous().view(torch.uint8).view(-1), non_blocking=True)
                sm.copy_(masks.contiguous().view(torch.uint8).view(-1), non_blocking=True)
                self.ev_staged[slot].record(self.cs)      # the producer may overwrite dets / masks from here on
            else:
                # pushed straight from the caller's tensors: they must stay untouched until done_event(slot)
                sd, sm = dets.contiguous().view(torch.uint8).view(-1), masks.contiguous().view(torch.uint8).view(-1)
            for dst_d, dst_m, dst_flag, ack_addr in self.push[slot]:
                if seq > 1:                                # that peer has read what I pushed into this slot last time
                    self._wait_flag(cs_h, ack_addr, seq - 1)
                dst_d.copy_(sd, non_blocking=True)
                dst_m.copy_(sm, non_blocking=True)
                dst_flag.copy_(val, non_blocking=True)     # stream order: after the payload
            for addr in self.wait_addr[slot]:              # every rank's payload has landed in my slot
                self._wait_flag(cs_h, addr, seq)
            self.ev_done[slot].record(self.cs)
        if stage:
            main.wait_event(self.ev_staged[slot])

    def _finish_ce(self, slot):
        """Views into the receive slot (no copy): valid until the slot is reused by the gather after next."""
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self.ev_done[slot])
        d, m = self.out_d[slot], self.out_m[slot]
        if not d.is_contiguous():
            d = d.contiguous()
        if not m.is_contiguous():
            m = m.contiguous()
        d = d.view(self.d_dtype).view((self.world * self.d_shape[0],) + self.d_shape[1:])
        m = m.view(self.m_dtype).view((self.world * self.m_shape[0],) + self.m_shape[1:])
        return d, m

    def _release_ce(self, slot):
        """The consumer is done with the slot (everything it enqueued on the current stream so far): every sender may
        push into it again."""
        val = self.seq_vals[4 * self.seq[slot]:4 * self.seq[slot] + 4]
        for dst in self.ack_dst[slot]:
            dst.copy_(val, non_blocking=True)

    # ---- public ---------------------------------------------------------------------------------
    def release(self, slot):
        """The caller has enqueued (on the current stream) every read of the results ``finish(slot)`` returned: tell the
        senders now that the slot may be overwritten, instead of at the next ``start()`` that reuses it.  With more
        than two slots this gives every sender its acknowledgement whole steps before it needs it, so a rank that runs
        late does not hold up the pushes of the others."""
        if self.world > 1 and self.transport == "ce" and self.seq[slot] > 0 and not self.released[slot]:
            self._release_ce(slot)
            self.released[slot] = True

    def done_event(self, slot):
        """CUDA event recorded when the gather started in `slot` is complete (copy-engine transport), else None."""
        return self.ev_done[slot] if (self.world > 1 and self.transport == "ce") else None

    def start(self, dets, masks, stage=True):
        """Begin the gather of this batch (asynchronous).  With ``stage`` (default) dets / masks may be overwritten by
        work enqueued on the current stream after this call returns (they are copied to a staging buffer first, and the
        current stream waits for that copy); ``stage=False`` pushes straight from the caller's tensors, which must then
        stay untouched until ``done_event(slot)`` (a producer that alternates between two output buffers)."""
        slot = self.count % self.slots
        self.count += 1
        if self.world == 1:
            self.outs[slot] = (dets, masks)
        elif self.transport == "ce":
            if self.seq[slot] > 0 and not self.released[slot]:
                self._release_ce(slot)                     # whatever read the previous contents was enqueued before
            self.released[slot] = False
            self._start_ce(dets, masks, slot, stage)
        else:
            if self.works[slot] is not None:
                for w in self.works[slot]:
                    w.wait()
            d, m, works = all_gather_frames(dets, masks, group=self.group, async_op=True)
            self.outs[slot], self.works[slot] = (d, m), works
        return slot

    def finish(self, slot):
        """Gathered (dets, masks) of the batch started in `slot`; the current stream waits for them."""
        if self.world > 1 and self.transport == "ce":
            return self._finish_ce(slot)
        if self.works[slot] is not None:
            for w in self.works[slot]:
                w.wait()
            self.works[slot] = None
        return self.outs[slot]


def run_clip(batches, step_fn, records_fn, group=None, device=None, transport="auto"):
    """Frame-sharded clip driver (reference twin: lib/core/test_engine.py:168-213, one process per range of frames,
    results merged at the end).  ``batches``: this rank's batches (every rank the same number of equally sized
    batches: pad the clip); ``step_fn(batch)`` runs the region pipeline on one batch (RegionPipeline.step on GPUs);
    ``records_fn(batch, out)`` -> (dets (F_r,D,6) fp32, masks (F_r,D,...) uint8) to exchange.  The gather of batch i
    runs behind the kernels of batch i+1.  Returns the clip's (dets, masks) with the frames of every batch in rank
    order: frame f of the clip = batch f // (world * F_r), rank (f // F_r) % world."""
    gather, started, d_all, m_all = None, [], [], []
    for batch in batches:
        out = step_fn(batch)
        dets, masks = records_fn(batch, out)
        if gather is None:
            gather = FrameGather(dets.shape, dets.dtype, masks.shape, masks.dtype, device or dets.device, group, transport)
        if len(started) == gather.slots:                   # the slot about to be reused: collect its result first
            d, m = gather.finish(started.pop(0))
            d_all.append(d.clone())
            m_all.append(m.clone())
        started.append(gather.start(dets, masks))
    while started:
        d, m = gather.finish(started.pop(0))
        d_all.append(d.clone())
        m_all.append(m.clone())
    if not d_all:
        return None, None
    return torch.cat(d_all), torch.cat(m_all)


class HostPipeline:
    """Host-buffer front end of RegionPipeline.step: pinned host tensors in, pinned host tensors out.

    Three CUDA streams (H2D, compute, D2H) and ``depth`` rotating slots, so that the upload of batch
    i+1, the kernels of batch i and the download of batch i-1 overlap (PCIe is full duplex and the
    copy engines run beside the SMs).  Every batch still pays its own H2D of all inputs and its own D2H
    of the results; nothing is cached across batches.

        hp = HostPipeline(RegionPipeline(cfg), frame_hw, im_scale, device)
        for batch in clip: hp.submit(batch)        # dict of pinned host tensors, see ``INPUT_KEYS``
        hp.synchronize(); out = hp.results(slot)
    """
    INPUT_KEYS = ("rpn", "im_info", "feats", "det_boxes", "det_cls", "det_masks")

    def __init__(self, pipe, frame_hw, im_scale, device, depth=2, masks="packed", select_class_channel=True):
        """``masks``: what comes back to the host for the pasted masks -- "packed" (default: 1 bit per pixel, written by
        the paste kernel itself, 8x fewer D2H bytes; ``unpack_mask_bits`` restores the reference's uint8 layout),
        "dense" (the reference's (B,D,H,W) uint8).  ``select_class_channel``: segm_results reads ONE of the K class
        channels of every detection's mask (masks[i, j], lib/core/test.py:826-829), so the (B,D,K,M,M) host tensor is
        gathered to (B,D,1,M,M) on the host before the upload (3 MB instead of 254 MB per 10 frames)."""
        if masks not in ("packed", "dense"):
            raise ValueError("masks must be 'packed' or 'dense'")
        self.masks, self.select_class_channel = masks, select_class_channel
        self.sel_host = [None] * depth      # pinned (B,D,1,M,M) staging of the selected channels
        self.pipe, self.frame_hw, self.im_scale, self.device, self.depth = pipe, frame_hw, im_scale, device, depth
        self.s_h2d, self.s_cmp, self.s_d2h = (torch.cuda.Stream(device) for _ in range(3))
        self.dev = [None] * depth          # device input slots (allocated on first use)
        self.host_out = [None] * depth     # pinned output slots
        self.dev_out = [None] * depth      # device results of the batch in flight in each slot
        self.ev_h2d = [torch.cuda.Event() for _ in range(depth)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_d2h = [torch.cuda.Event() for _ in range(depth)]
        self.count = 0

    @staticmethod
    def _like(x, device):
        if isinstance(x, dict):
            return {k: HostPipeline._like(v, device) for k, v in x.items()}
        if isinstance(x, (tuple, list)):
            return tuple(HostPipeline._like(v, device) for v in x)
        cl = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        return torch.empty(x.shape, dtype=x.dtype, device=device,
                           memory_format=torch.channels_last if cl else torch.contiguous_format)

    @staticmethod
    def _copy(dst, src):
        if isinstance(src, dict):
            for k in src:
                HostPipeline._copy(dst[k], src[k])
        elif isinstance(src, (tuple, list)):
            for d, s in zip(dst, src):
                HostPipeline._copy(d, s)
        else:
            dst.copy_(src, non_blocking=True)

    def submit(self, host):
        """Enqueue one batch (H2D -> step -> D2H) without blocking the host; returns the slot index."""
        slot = self.count % self.depth
        self.count += 1
        host = {k: host[k] for k in self.INPUT_KEYS}
        m = host["det_masks"]
        if self.select_class_channel and m.shape[2] > 1 and self.pipe.cfg.mrcnn_cls_specific_mask:
            # host-side gather of masks[b, d, cls[b, d]] into a pinned staging buffer (part of the step's host work)
            if self.sel_host[slot] is None:
                self.sel_host[slot] = torch.empty((m.shape[0], m.shape[1], 1) + tuple(m.shape[3:]), dtype=m.dtype).pin_memory()
            self.ev_h2d[slot].synchronize()                 # the upload that last read this staging buffer is done
            idx = host["det_cls"].long().view(m.shape[0], m.shape[1], 1, 1, 1).expand(-1, -1, 1, m.shape[3], m.shape[4])
            torch.gather(m, 2, idx, out=self.sel_host[slot])
            host["det_masks"] = self.sel_host[slot]
        if self.dev[slot] is None:
            self.dev[slot] = self._like(host, self.device)
        d = self.dev[slot]
        with torch.cuda.stream(self.s_h2d):
            self.s_h2d.wait_event(self.ev_cmp[slot])        # the step that last read this slot is done
            self._copy(d, host)
            self.ev_h2d[slot].record(self.s_h2d)
        keep_mode = self.pipe.packed_masks
        if self.masks == "packed" and not keep_mode:
            self.pipe.packed_masks = True                   # the paste kernel writes the 1-bit copy itself
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(self.ev_h2d[slot])
            out = self.pipe.step(d["rpn"], d["im_info"], d["feats"], d["det_boxes"], d["det_cls"], d["det_masks"],
                                 self.frame_hw, self.im_scale)
            self.ev_cmp[slot].record(self.s_cmp)
        self.pipe.packed_masks = keep_mode
        self.dev_out[slot] = out
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(self.ev_cmp[slot])
            if self.host_out[slot] is None:
                keys = ("rois", "roi_count", "masks" if self.masks == "dense" else "masks_packed")
                self.host_out[slot] = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in keys}
            for k, h in self.host_out[slot].items():
                h.copy_(out[k], non_blocking=True)
                out[k].record_stream(self.s_d2h)
            self.ev_d2h[slot].record(self.s_d2h)
        return slot

    def synchronize(self):
        for s in (self.s_h2d, self.s_cmp, self.s_d2h):
            s.synchronize()

    def results(self, slot):
        """Host results of the batch last submitted to `slot` (blocks until its D2H finished); the RoIAlign
        blobs stay on the device for the heads that consume them."""
        self.ev_d2h[slot].synchronize()
        out = dict(self.host_out[slot])
        out["box_feats"], out["mask_feats"] = self.dev_out[slot]["box_feats"], self.dev_out[slot]["mask_feats"]
        return out
