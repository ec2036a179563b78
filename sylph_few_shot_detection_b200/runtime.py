"""Thin Python handle over the C-ABI context: torch is used only for device memory and the current CUDA stream."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char, c_double, c_float, c_int, c_int64, c_void_p
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (CG_MAX_TOWER, CODE_STRIDE, DET_STRIDE, IPC_HANDLE_BYTES, LOSS_SUMS, NUM_LEVELS, SLOT_QUERY, SLOT_SUPPORT, CodegenTensors,
                   LossConfig, ModelConfig, TowerTensors)


def model_config_from_cfg(cfg) -> ModelConfig:
    """Collect the config keys the hot path reads into the C struct."""
    F, G = cfg.MODEL.FCOS, cfg.MODEL.META_LEARN.CODE_GENERATOR
    if G.NAME not in ("CodeGenerator", "CodeGeneratorHead", "ROIEncoder"):
        raise NotImplementedError(f"code generator {G.NAME!r} is not implemented on the B200 path")
    roi_encoder = G.NAME == "ROIEncoder"
    if F.NORM != "GN" or F.USE_DEFORMABLE or F.NUM_SHARE_CONVS != 0 or list(F.IN_FEATURES) != ["p3", "p4", "p5", "p6", "p7"]:
        raise NotImplementedError("FCOS head variant outside the shipped Meta-FCOS configs")
    if list(F.FPN_STRIDES) != [8, 16, 32, 64, 128] or int(F.TOP_LEVELS) != 2:
        raise NotImplementedError("only the p3..p7 pyramid (strides 8..128) is implemented")
    if cfg.MODEL.RESNETS.NORM != "FrozenBN" or not cfg.MODEL.RESNETS.STRIDE_IN_1X1:
        raise NotImplementedError("backbone must use FrozenBN and STRIDE_IN_1X1 (inference path)")
    episodic = bool(cfg.MODEL.META_LEARN.EPISODIC_LEARNING)
    if not episodic:
        roi_encoder = False              # base detector: no code generator is built, its config keys are not read
    elif roi_encoder:
        T, E, H = G.TOKENIZER, G.TRANSFORMER_ENCODER, G.HEAD
        if int(T.CONV_DIM) != 256 or int(T.FC_DIM) != 256 or int(H.OUTPUT_DIM) != 256 or T.NORM != "GN":
            raise NotImplementedError("ROIEncoder: TOKENIZER.CONV_DIM / FC_DIM and HEAD.OUTPUT_DIM must be 256, TOKENIZER.NORM 'GN'")
        if int(T.NUM_CONV) < 1 or int(T.NUM_FC) < 1 or int(H.NUM_FC) < 1 or int(H.FC_DIM) > 1024 or 256 % int(E.HEADS):
            raise NotImplementedError("ROIEncoder: unsupported tokenizer / head depth or width")
    if episodic and not roi_encoder:
        for layer in G.TOWER_LAYERS:
            if list(layer) != ["GN", "ReLU"]:
                raise NotImplementedError("CODE_GENERATOR.TOWER_LAYERS entries must be ['GN', 'ReLU']")
        if list(G.CLS_LAYER) != ["", "", 1]:
            raise NotImplementedError("CODE_GENERATOR.CLS_LAYER must be ['', '', 1]")
        if len(G.WEIGHT_LAYER) not in (0, 3) or (len(G.WEIGHT_LAYER) == 3 and G.WEIGHT_LAYER[0] != ""):
            raise NotImplementedError("CODE_GENERATOR.WEIGHT_LAYER must be [] or ['', <act>, <pool>] (no norm layer on the 1-channel head)")
        if len(G.SCALE_LAYER) or G.COMPRESS_CODE_W_MAX or G.ROI_BOX.FPN_MULTILEVEL_FEATURE:
            raise NotImplementedError("SCALE_LAYER / COMPRESS_CODE_W_MAX / multilevel ROI are not implemented")
    if episodic and (G.ROI_BOX.POOLER_TYPE != "ROIAlignV2" or int(G.ROI_BOX.POOLER_RESOLUTION) != 7):
        raise NotImplementedError("ROI pooler must be ROIAlignV2 at 7x7")
    # MODEL.PROPOSAL_GENERATOR.OWD only selects the evaluator (meta_fcos_runner.py:121-126: COCO_OWD_Evaluator, class-agnostic
    # matching); nothing under sylph/modeling reads it, so the forward path is the same either way.
    quality = sorted(F.BOX_QUALITY)
    mc = ModelConfig()
    mc.resnet_depth = int(cfg.MODEL.RESNETS.DEPTH)
    mc.num_cls_convs, mc.num_box_convs = int(F.NUM_CLS_CONVS), int(F.NUM_BOX_CONVS)
    mc.use_scale = int(bool(F.USE_SCALE))
    mc.thresh_with_ctr = int(bool(F.THRESH_WITH_CTR))
    mc.box_quality = (1 if "ctrness" in quality else 0) | (2 if "iou" in quality else 0)
    mc.pre_nms_topk, mc.post_nms_topk = int(F.PRE_NMS_TOPK_TEST), int(F.POST_NMS_TOPK_TEST)
    mc.inference_thresh, mc.nms_thresh, mc.prior_prob = float(F.INFERENCE_TH_TEST), float(F.NMS_TH), float(F.PRIOR_PROB)
    for i in range(3):
        mc.pixel_mean[i] = float(cfg.MODEL.PIXEL_MEAN[i])
        mc.pixel_std[i] = float(cfg.MODEL.PIXEL_STD[i])
    mc.cg_tower_layers = len(G.TOWER_LAYERS)
    mc.cg_post_norm = int(G.POST_NORM == "GN")
    if episodic and G.POST_NORM not in ("", "GN"):
        raise NotImplementedError("POST_NORM must be '' or 'GN'")
    mc.cg_conv_l2_norm = int(bool(G.CONV_L2_NORM))
    mc.cg_bias_layer = int(len(G.BIAS_LAYER) == 3)
    mc.cg_weight_layer = int(episodic and not roi_encoder and len(G.WEIGHT_LAYER) == 3)
    mc.cg_bias_l2_norm = int(bool(G.BIAS_L2_NORM))
    mc.cg_use_bias = int(bool(G.USE_BIAS))
    mc.cg_has_conv_scale = int(bool(G.USE_WEIGHT_SCALE and (G.CONV_L2_NORM or G.POST_NORM != "")))
    # 0 = CodeGenerator, 1 = ROIEncoder, 2 = none (base detector: MODEL.META_LEARN.EPISODIC_LEARNING False, cls_logits classifier)
    mc.generator = int(roi_encoder) if cfg.MODEL.META_LEARN.EPISODIC_LEARNING else 2
    if roi_encoder:
        mc.re_tok_convs, mc.re_tok_fcs = int(G.TOKENIZER.NUM_CONV), int(G.TOKENIZER.NUM_FC)
        mc.re_layers = int(G.TRANSFORMER_ENCODER.LAYERS)
        mc.re_head_fcs, mc.re_head_dim = int(G.HEAD.NUM_FC), int(G.HEAD.FC_DIM)
    return mc


def loss_config_from_cfg(cfg) -> LossConfig:
    """The keys FCOSOutputs._init_fcos / _init_code_generator read for the training losses
    (sylph/modeling/meta_fcos/fcos_outputs.py:70-123)."""
    F, G = cfg.MODEL.FCOS, cfg.MODEL.META_LEARN.CODE_GENERATOR
    kinds = {"iou": 0, "linear_iou": 1, "giou": 2}
    if F.LOC_LOSS_TYPE not in kinds:
        raise NotImplementedError(F.LOC_LOSS_TYPE)                      # IOULoss.forward, iou_loss.py:80-81
    if G.NAME == "CodeGenerator" and G.BOX_ON:
        raise NotImplementedError("CODE_GENERATOR.BOX_ON (per-class box regression) is not implemented")
    if float(G.DISTILLATION_LOSS_WEIGHT) != 0.0:
        raise NotImplementedError("DISTILLATION_LOSS_WEIGHT > 0 is not implemented (0.0 in the shipped configs)")
    if G.CONTRASTIVE_LOSS not in ("", None):
        raise NotImplementedError("CONTRASTIVE_LOSS (snnl) is not implemented ('' in the shipped configs)")
    if len(F.SIZES_OF_INTEREST) != 4:
        raise NotImplementedError("SIZES_OF_INTEREST must have 4 entries (5 FPN levels)")
    lc = LossConfig()
    lc.focal_alpha, lc.focal_gamma = float(F.LOSS_ALPHA), float(F.LOSS_GAMMA)
    lc.center_sample, lc.pos_radius = int(bool(F.CENTER_SAMPLE)), float(F.POS_RADIUS)
    lc.loc_loss_type = kinds[F.LOC_LOSS_TYPE]
    for i in range(4):
        lc.sizes_of_interest[i] = int(F.SIZES_OF_INTEREST[i])
    return lc


class Engine:
    """One context per device.  Raises RuntimeError with the library's message on any non-zero status."""

    def __init__(self, cfg, device: int = 0, precision: Optional[str] = None):
        """`precision`: "exact" (split-fp16 operands, three tensor-core products per multiply: fp32-level agreement
        with the reference) or "fast" (single fp16 operands).  None = $SYLPH_PRECISION, else "exact"."""
        if not torch.cuda.is_available():
            raise RuntimeError("sylph_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.cfg = cfg
        self.mc = model_config_from_cfg(cfg)
        self.post_nms_topk = self.mc.post_nms_topk
        if precision is None:
            precision = os.environ.get("SYLPH_PRECISION", "exact")
        if precision not in ("exact", "fast"):
            raise ValueError(f"precision must be 'exact' or 'fast', got {precision!r}")
        h = c_void_p()
        rc = self.lib.sylph_create(byref(h), device, byref(self.mc))
        if rc != 0 or not h:
            raise RuntimeError(f"sylph_create failed with status {rc} (needs an sm_100 device)")
        self.h = h
        self._check(self.lib.sylph_set_precision(self.h, 1 if precision == "exact" else 0))
        self.precision = precision
        self._loaded = False

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.sylph_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError("sylph_b200: " + self.lib.sylph_last_error(self.h).decode())

    @staticmethod
    def _stream() -> c_void_p:
        return c_void_p(torch.cuda.current_stream().cuda_stream)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state: Dict[str, torch.Tensor]) -> None:
        for k, v in state.items():
            t = v.detach().to("cpu", torch.float32).contiguous()
            shape = (c_int64 * max(t.dim(), 1))(*t.shape)
            self._check(self.lib.sylph_load_tensor(self.h, k.encode(), c_void_p(t.data_ptr()), shape, t.dim()))
        self._check(self.lib.sylph_finalize_weights(self.h))
        self._loaded = True

    # ------------------------------------------------------------------ features
    def extract_features(self, slot: int, images: Sequence[torch.Tensor]) -> None:
        u8 = all(im.dtype == torch.uint8 for im in images)
        imgs = [im.to(self.device, torch.uint8 if u8 else torch.float32, non_blocking=True).contiguous() for im in images]
        n = len(imgs)
        ptrs = (c_void_p * n)(*[im.data_ptr() for im in imgs])
        hs = (c_int * n)(*[int(im.shape[-2]) for im in imgs])
        ws = (c_int * n)(*[int(im.shape[-1]) for im in imgs])
        fn = self.lib.sylph_extract_features_u8 if u8 else self.lib.sylph_extract_features
        self._check(fn(self.h, slot, n, ptrs, hs, ws, self._stream()))
        self._keep = imgs  # keep inputs alive until the stream work is done

    def extract_features_multi(self, groups: Sequence[Tuple[int, Sequence[torch.Tensor]]]) -> None:
        """[(slot, images), ...] through one trunk pass when the groups pad to the same size (sylph_extract_features_multi)."""
        flat = [im for _, ims in groups for im in ims]
        u8 = all(im.dtype == torch.uint8 for im in flat)
        imgs = [im.to(self.device, torch.uint8 if u8 else torch.float32, non_blocking=True).contiguous() for im in flat]
        n = len(imgs)
        slots = (c_int * len(groups))(*[int(s) for s, _ in groups])
        counts = (c_int * len(groups))(*[len(ims) for _, ims in groups])
        ptrs = (c_void_p * n)(*[im.data_ptr() for im in imgs])
        hs = (c_int * n)(*[int(im.shape[-2]) for im in imgs])
        ws = (c_int * n)(*[int(im.shape[-1]) for im in imgs])
        self._check(self.lib.sylph_extract_features_multi(self.h, len(groups), slots, counts, ptrs, int(u8), hs, ws, self._stream()))
        self._keep = imgs

    def extract_features_normalized(self, slot: int, batch: torch.Tensor) -> None:
        """The backbone on an already normalised, zero-padded (N, 3, H, W) batch (what the reference hands to
        `self.backbone(images.tensor)`)."""
        x = batch.to(self.device, torch.float32).contiguous()
        assert x.dim() == 4 and x.shape[1] == 3, "expected a (N, 3, H, W) batch"
        self._check(self.lib.sylph_extract_features_normalized(self.h, slot, x.shape[0], c_void_p(x.data_ptr()),
                                                               int(x.shape[2]), int(x.shape[3]), self._stream()))
        x.record_stream(torch.cuda.current_stream())

    def import_features(self, slot: int, features: Sequence[torch.Tensor], padded_hw: Tuple[int, int],
                        image_sizes: Optional[Sequence[Tuple[int, int]]] = None) -> None:
        """`image_sizes`: the un-padded (h, w) of every image (ImageList.image_sizes); detection boxes are scaled by
        out_size / image_size, so they must be known for proposals to come back un-scaled."""
        feats = [f.to(self.device, torch.float32).contiguous() for f in features]
        assert len(feats) == NUM_LEVELS and all(f.shape[1] == 256 for f in feats)
        n = feats[0].shape[0]
        ptrs = (c_void_p * NUM_LEVELS)(*[f.data_ptr() for f in feats])
        lh = (c_int * NUM_LEVELS)(*[int(f.shape[2]) for f in feats])
        lw = (c_int * NUM_LEVELS)(*[int(f.shape[3]) for f in feats])
        self._check(self.lib.sylph_import_features(self.h, slot, n, int(padded_hw[0]), int(padded_hw[1]), ptrs, lh, lw,
                                                   self._stream()))
        if image_sizes is not None:
            hs = (c_int * n)(*[int(s[0]) for s in image_sizes])
            ws = (c_int * n)(*[int(s[1]) for s in image_sizes])
            self._check(self.lib.sylph_set_image_sizes(self.h, slot, n, hs, ws))
        self._keep = feats

    def feature_shape(self, slot: int):
        n, ph, pw = c_int(), c_int(), c_int()
        lh, lw = (c_int * NUM_LEVELS)(), (c_int * NUM_LEVELS)()
        self._check(self.lib.sylph_feature_shape(self.h, slot, byref(n), byref(ph), byref(pw), lh, lw))
        return n.value, ph.value, pw.value, list(lh), list(lw)

    def export_features(self, slot: int, level: int) -> torch.Tensor:
        n, _, _, lh, lw = self.feature_shape(slot)
        out = torch.empty((n, 256, lh[level], lw[level]), device=self.device, dtype=torch.float32)
        self._check(self.lib.sylph_export_features(self.h, slot, level, c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------ class codes
    def generate_codes(self, slot: int, boxes: torch.Tensor, roi_image: Sequence[int], class_offsets: Sequence[int],
                       want_levels: bool = False):
        # host boxes are staged through the pinned ring; a CUDA tensor is read in place (static buffer of a CUDA graph)
        boxes = boxes.detach().to(torch.float32).contiguous()
        n_rois, n_classes = boxes.shape[0], len(class_offsets) - 1
        codes = torch.empty((n_classes, CODE_STRIDE), device=self.device, dtype=torch.float32)
        levels = torch.empty((n_rois,), device=self.device, dtype=torch.int64) if want_levels else None
        ri = (c_int * n_rois)(*[int(i) for i in roi_image])
        co = (c_int * (n_classes + 1))(*[int(i) for i in class_offsets])
        self._check(self.lib.sylph_generate_codes(
            self.h, slot, n_rois, ctypes.cast(c_void_p(boxes.data_ptr()), POINTER(c_float)), ri, n_classes, co,
            c_void_p(codes.data_ptr()), c_void_p(levels.data_ptr()) if want_levels else None, self._stream()))
        self._keep_boxes = boxes   # keep the argument alive until the stream work has run
        return (codes, levels) if want_levels else codes

    def export_roi_features(self, n_rois: int) -> torch.Tensor:
        out = torch.empty((n_rois, 256, 7, 7), device=self.device, dtype=torch.float32)
        self._check(self.lib.sylph_export_roi_features(self.h, c_void_p(out.data_ptr()), self._stream()))
        return out

    def normalize_codes(self, raw: torch.Tensor) -> torch.Tensor:
        raw = raw.to(self.device, torch.float32).contiguous()
        out = torch.empty_like(raw)
        self._check(self.lib.sylph_normalize_codes(self.h, c_void_p(raw.data_ptr()), c_void_p(out.data_ptr()),
                                                   raw.shape[0], self._stream()))
        return out

    # ------------------------------------------------------------------ class-code exchange over NVLink peer memory
    def exchange_setup(self, group=None, max_classes: int = 2048) -> None:
        """Create this rank's exchange buffer, swap the CUDA IPC handles of all ranks of the group (ONE all-gather of
        64 bytes, at setup only) and map the peers' buffers (sylph_exchange_create / _connect).  All ranks must be GPUs
        of one box.  Idempotent for the same group size."""
        import torch.distributed as dist
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        if getattr(self, "_xch", None) == (world, rank, max_classes):
            return
        if getattr(self, "_xch", None) is not None:
            self.exchange_teardown(group)
        handle = (ctypes.c_uint8 * IPC_HANDLE_BYTES)()
        self._check(self.lib.sylph_exchange_create(self.h, world, rank, int(max_classes), handle))
        handles = None
        if world > 1:
            dev = self.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            everyone = torch.empty((world * IPC_HANDLE_BYTES,), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(everyone, mine, group=group)
            handles = (ctypes.c_uint8 * (world * IPC_HANDLE_BYTES))(*everyone.cpu().tolist())
        self._check(self.lib.sylph_exchange_connect(self.h, handles))
        self._xch = (world, rank, max_classes)

    def normalize_codes_exchange(self, raw_local: Optional[torch.Tensor], class_offset: int, n_total: int) -> torch.Tensor:
        """Normalise this rank's (n_local, 257) raw codes = classes [class_offset, class_offset + n_local) of the episode,
        deliver them to every rank and return the (n_total, 257) normalised codes of ALL classes
        (sylph_normalize_codes_exchange: two kernels, the stores travel over NVLink; no NCCL call, no host sync)."""
        assert getattr(self, "_xch", None) is not None, "call exchange_setup(group) first"
        n_local = 0 if raw_local is None else int(raw_local.shape[0])
        raw = raw_local.to(self.device, torch.float32).contiguous() if n_local else None
        out = torch.empty((n_total, CODE_STRIDE), device=self.device, dtype=torch.float32)
        self._check(self.lib.sylph_normalize_codes_exchange(self.h, c_void_p(raw.data_ptr()) if n_local else None, n_local,
                                                            int(class_offset), int(n_total), c_void_p(out.data_ptr()),
                                                            self._stream()))
        self._keep_raw = raw
        return out

    def exchange_poll(self) -> None:
        """Raise RuntimeError if an exchange of this context gave up waiting for its peers (non-blocking: reads the pinned
        copy of the flag, which is current for every episode whose results have been synchronised on)."""
        self._check(self.lib.sylph_exchange_poll(self.h))

    def exchange_status(self) -> Tuple[bool, int]:
        """(timed_out, rows_arrived) -- synchronises the device."""
        flag, rows = c_int(), c_int64()
        self._check(self.lib.sylph_exchange_status(self.h, byref(flag), byref(rows)))
        return bool(flag.value), int(rows.value)

    def exchange_teardown(self, group=None) -> None:
        """Unmap the peers and free the buffer; the barrier keeps a fast rank from freeing memory a peer still writes."""
        import torch.distributed as dist
        if getattr(self, "_xch", None) is None:
            return
        if self._xch[0] > 1 and dist.is_initialized():
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)
        self.lib.sylph_exchange_destroy(self.h)
        self._xch = None

    # ------------------------------------------------------------------ base-class accumulation / reduction
    def accumulate_codes(self, chunk_codes: torch.Tensor, chunk_class: Sequence[int], chunk_weight: Sequence[float],
                         acc: torch.Tensor) -> torch.Tensor:
        """acc[class_k] += chunk_codes[k] * float32(chunk_weight[k]) in chunk order (in place on the device)."""
        chunk_codes = chunk_codes.to(self.device, torch.float32).contiguous()
        n = chunk_codes.shape[0]
        assert acc.is_cuda and acc.dtype == torch.float32 and acc.is_contiguous() and acc.shape[1] == CODE_STRIDE
        cc = (c_int * n)(*[int(v) for v in chunk_class])
        cw = (c_float * n)(*[float(v) for v in chunk_weight])
        self._check(self.lib.sylph_accumulate_codes(self.h, c_void_p(chunk_codes.data_ptr()), n, cc, cw,
                                                    c_void_p(acc.data_ptr()), acc.shape[0], self._stream()))
        return acc

    def reduce_codes(self, parts: torch.Tensor, divisor: Sequence[float]) -> torch.Tensor:
        """parts (n_parts, n_classes, 257) -> (n_classes, 257): ordered sum over parts, then / divisor where != 0."""
        parts = parts.to(self.device, torch.float32).contiguous()
        n_parts, n_classes = parts.shape[0], parts.shape[1]
        out = torch.empty((n_classes, CODE_STRIDE), device=self.device, dtype=torch.float32)
        dv = (c_float * max(n_classes, 1))(*[float(v) for v in divisor])
        self._check(self.lib.sylph_reduce_codes(self.h, c_void_p(parts.data_ptr()), n_parts, n_classes, dv,
                                                c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------ detection
    def detect(self, slot: int, codes: torch.Tensor, out_sizes: Optional[Sequence[Tuple[int, int]]] = None,
               max_dets: Optional[int] = None, codes_ready: Optional["torch.cuda.Event"] = None):
        """`codes_ready`: event recorded on the stream that produced `codes` (sylph_detect_after): the towers are
        enqueued on the current stream first, which waits for the event only before the code-conditioned classifier."""
        if codes_ready is None:
            codes = codes.to(self.device, torch.float32).contiguous()
        else:
            assert codes.is_cuda and codes.dtype == torch.float32 and codes.is_contiguous()
        n = self.feature_shape(slot)[0]
        max_dets = max_dets or min(1024, max(self.post_nms_topk * 2, 128))
        dets = torch.zeros((n, max_dets, DET_STRIDE), device=self.device, dtype=torch.float32)
        counts = torch.zeros((n,), device=self.device, dtype=torch.int32)
        sizes = None
        if out_sizes is not None:
            flat = [int(v) for hw in out_sizes for v in hw]
            sizes = (c_int * len(flat))(*flat)
        if codes_ready is None:
            self._check(self.lib.sylph_detect(self.h, slot, c_void_p(codes.data_ptr()), codes.shape[0], sizes,
                                              c_void_p(dets.data_ptr()), c_void_p(counts.data_ptr()), max_dets, self._stream()))
        else:
            self._check(self.lib.sylph_detect_after(self.h, slot, c_void_p(codes.data_ptr()), codes.shape[0], sizes,
                                                    c_void_p(dets.data_ptr()), c_void_p(counts.data_ptr()), max_dets,
                                                    c_void_p(codes_ready.cuda_event), self._stream()))
            codes.record_stream(torch.cuda.current_stream())
        return dets, counts

    def detect_poll(self) -> None:
        """Raises if a finished detect call dropped candidates (list overflow); call after synchronising on its results."""
        self._check(self.lib.sylph_detect_poll(self.h))

    def side_stream(self) -> "torch.cuda.Stream":
        """Context-owned second stream: code generation runs here while the code-independent FCOS towers run on the
        caller's stream (generate_and_detect, runner.run_episode)."""
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    @staticmethod
    def overlap_enabled() -> bool:
        # opt-in: measured neutral on one GPU (14.13 vs 14.15 ms per episode: the small code-generation kernels take the
        # SMs they run on away from the persistent tower kernels), see DESIGN.md section 4
        return os.environ.get("SYLPH_OVERLAP_CODEGEN", "0") == "1"

    def generate_and_detect(self, sup_slot: int, qry_slot: int, boxes: torch.Tensor, roi_image: Sequence[int],
                            class_offsets: Sequence[int], normalize: bool = True,
                            out_sizes: Optional[Sequence[Tuple[int, int]]] = None, max_dets: Optional[int] = None,
                            overlap: Optional[bool] = None):
        """Code generation (+ normalisation) from the support slot and detection on the query slot, both slots already
        filled.  With `overlap` the ROIAlign / code-tower / mean / normalise kernels (small grids, latency-bound) run on
        the side stream while the current stream runs the class / box towers; the streams join right before the
        code-conditioned classifier (sylph_detect_after).  Returns ((dets, counts), codes)."""
        if overlap is None:
            overlap = self.overlap_enabled()
        if not overlap:
            raw = self.generate_codes(sup_slot, boxes, roi_image, class_offsets)
            codes = self.normalize_codes(raw) if normalize else raw
            return self.detect(qry_slot, codes, out_sizes, max_dets), codes
        cur, side = torch.cuda.current_stream(), self.side_stream()
        side.wait_stream(cur)                      # the feature slots were written on the current stream
        with torch.cuda.stream(side):
            raw = self.generate_codes(sup_slot, boxes, roi_image, class_offsets)
            codes = self.normalize_codes(raw) if normalize else raw
            ready = torch.cuda.Event()
            ready.record(side)
        return self.detect(qry_slot, codes, out_sizes, max_dets, codes_ready=ready), codes

    def export_head_output(self, which: int, level: int, slot: int, n_classes: int) -> torch.Tensor:
        n, _, _, lh, lw = self.feature_shape(slot)
        ch = {0: n_classes, 1: 4, 2: 1, 3: 1}[which]
        out = torch.empty((n, ch, lh[level], lw[level]), device=self.device, dtype=torch.float32)
        self._check(self.lib.sylph_export_head_output(self.h, which, level, c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------ training forward (losses only)
    def fcos_loss_sums(self, slot: int, codes: torch.Tensor, support_targets: Sequence[int], gt_boxes: torch.Tensor,
                       gt_classes: torch.Tensor, gt_offsets: Sequence[int], want_targets: bool = False):
        """Head + ground-truth assignment + per-rank loss sums (sylph_fcos_loss_sums).  Returns the (5,) float64 sums
        on the device and, with want_targets, (labels, target_inds, reg_targets) in level-first order."""
        codes = codes.to(self.device, torch.float32).contiguous()
        n, _, _, lh, lw = self.feature_shape(slot)
        total = n * sum(h * w for h, w in zip(lh, lw))
        n_cls = codes.shape[0]
        gt_boxes = gt_boxes.detach().to("cpu", torch.float32).reshape(-1, 4).contiguous()
        gt_classes = gt_classes.detach().to("cpu", torch.int64).reshape(-1).contiguous()
        n_gt = gt_boxes.shape[0]
        assert gt_classes.numel() == n_gt and len(gt_offsets) == n + 1 and len(support_targets) == n_cls
        st = (c_int64 * n_cls)(*[int(t) for t in support_targets])
        off = (c_int * (n + 1))(*[int(o) for o in gt_offsets])
        sums = torch.empty((LOSS_SUMS,), device=self.device, dtype=torch.float64)
        labels = inds = regs = None
        if want_targets:
            labels = torch.empty((total,), device=self.device, dtype=torch.int64)
            inds = torch.empty((total,), device=self.device, dtype=torch.int64)
            regs = torch.empty((total, 4), device=self.device, dtype=torch.float32)
        if not hasattr(self, "_lc"):
            self._lc = loss_config_from_cfg(self.cfg)
        ptr = lambda t: c_void_p(t.data_ptr()) if t is not None else None
        self._check(self.lib.sylph_fcos_loss_sums(
            self.h, slot, ptr(codes), n_cls, st, byref(self._lc), n_gt,
            ctypes.cast(c_void_p(gt_boxes.data_ptr()), POINTER(c_float)) if n_gt else None,
            ctypes.cast(c_void_p(gt_classes.data_ptr()), POINTER(c_int64)) if n_gt else None, off,
            ptr(sums), ptr(labels), ptr(inds), ptr(regs), self._stream()))
        return (sums, (labels, inds, regs)) if want_targets else sums

    def fcos_loss_finalize(self, sums: torch.Tensor, global_pos_ctr: Optional[torch.Tensor] = None, world_size: int = 1) -> torch.Tensor:
        """(loss_fcos_cls, loss_fcos_loc, loss_fcos_ctr) as a (3,) fp32 device tensor (sylph_fcos_loss_finalize)."""
        out = torch.empty((3,), device=self.device, dtype=torch.float32)
        if global_pos_ctr is not None:
            assert global_pos_ctr.is_cuda and global_pos_ctr.dtype == torch.float64 and global_pos_ctr.numel() == 2
        self._check(self.lib.sylph_fcos_loss_finalize(
            self.h, c_void_p(sums.data_ptr()), c_void_p(global_pos_ctr.data_ptr()) if global_pos_ctr is not None else None,
            int(world_size), c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------ training backward (code generator)
    CG_PREFIX = "code_generator.code_generator_head."

    def fcos_cls_loss_backward(self, slot: int, n_classes: int, support_targets: Sequence[int], labels: torch.Tensor,
                               sums: torch.Tensor, global_pos_ctr: Optional[torch.Tensor] = None, world_size: int = 1,
                               grad_loss: Optional[torch.Tensor] = None) -> torch.Tensor:
        """d loss_fcos_cls / d FINAL codes, (n_classes, 257) on the device (sylph_fcos_cls_loss_backward); call right after
        fcos_loss_sums(..., want_targets=True) on the same slot: `labels` and `sums` are that call's outputs."""
        assert labels.is_cuda and labels.dtype == torch.int64 and sums.is_cuda and sums.dtype == torch.float64
        st = (c_int64 * n_classes)(*[int(t) for t in support_targets])
        out = torch.empty((n_classes, CODE_STRIDE), device=self.device, dtype=torch.float32)
        if grad_loss is not None:
            grad_loss = grad_loss.detach().to(self.device, torch.float32).reshape(1).contiguous()
        if not hasattr(self, "_lc"):
            self._lc = loss_config_from_cfg(self.cfg)
        self._check(self.lib.sylph_fcos_cls_loss_backward(
            self.h, slot, n_classes, st, byref(self._lc), c_void_p(labels.data_ptr()), c_void_p(sums.data_ptr()),
            c_void_p(global_pos_ctr.data_ptr()) if global_pos_ctr is not None else None, int(world_size),
            c_void_p(grad_loss.data_ptr()) if grad_loss is not None else None, c_void_p(out.data_ptr()), self._stream()))
        return out

    def _codegen_tensor_struct(self, tensors: Dict[str, torch.Tensor]) -> CodegenTensors:
        """state_dict-keyed device fp32 tensors -> sylph_codegen_tensors (missing keys stay NULL)."""
        ct = CodegenTensors()
        def ptr(name):
            t = tensors.get(self.CG_PREFIX + name)
            if t is None:
                return None
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), name
            return t.data_ptr()
        for i in range(min(int(self.mc.cg_tower_layers), CG_MAX_TOWER)):
            ct.tower_w[i] = ptr(f"support_set_shared_tower.{3 * i}.weight")
            ct.tower_b[i] = ptr(f"support_set_shared_tower.{3 * i}.bias")
            ct.tower_gn_w[i] = ptr(f"support_set_shared_tower.{3 * i + 1}.weight")
            ct.tower_gn_b[i] = ptr(f"support_set_shared_tower.{3 * i + 1}.bias")
        ct.cls_w, ct.cls_b = ptr("support_set_cls_conv.0.weight"), ptr("support_set_cls_conv.0.bias")
        ct.bias_w, ct.bias_b = ptr("support_set_cls_bias.0.weight"), ptr("support_set_cls_bias.0.bias")
        ct.post_norm_w, ct.post_norm_b = ptr("post_norm.weight"), ptr("post_norm.bias")
        ct.conv_scale, ct.bias_scale = ptr("conv_scale.scale"), ptr("bias_scale.scale")
        return ct

    def codegen_backward(self, class_offsets: Sequence[int], raw_codes: torch.Tensor, grad_codes: torch.Tensor,
                         params: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Gradients of the code generator's tensors for the ROIs of the last generate_codes call (sylph_codegen_backward).
        `params`: {state_dict key: device fp32 tensor}; returns {key: gradient} for the same keys."""
        n_classes = len(class_offsets) - 1
        n_rois = int(class_offsets[-1])
        raw_codes = raw_codes.to(self.device, torch.float32).contiguous()
        grad_codes = grad_codes.to(self.device, torch.float32).contiguous()
        assert raw_codes.shape == (n_classes, CODE_STRIDE) and grad_codes.shape == (n_classes, CODE_STRIDE)
        params = {k: v for k, v in params.items() if k.startswith("code_generator.")}
        grads = {k: torch.zeros_like(v) for k, v in params.items()}
        co = (c_int * (n_classes + 1))(*[int(i) for i in class_offsets])
        p_struct, g_struct = self._codegen_tensor_struct(params), self._codegen_tensor_struct(grads)
        self._check(self.lib.sylph_codegen_backward(self.h, n_rois, n_classes, co, c_void_p(raw_codes.data_ptr()),
                                                    c_void_p(grad_codes.data_ptr()), byref(p_struct), byref(g_struct), self._stream()))
        self._keep_bwd = (raw_codes, grad_codes, params)   # alive until the stream work has run
        return grads

    def update_code_generator(self, state: Dict[str, torch.Tensor]) -> None:
        """Re-prepare the code generator's weights from `state` (every `code_generator.*` tensor) after an optimiser step."""
        for k, v in state.items():
            if not k.startswith("code_generator."):
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            shape = (c_int64 * max(t.dim(), 1))(*t.shape)
            self._check(self.lib.sylph_load_tensor(self.h, k.encode(), c_void_p(t.data_ptr()), shape, t.dim()))
        self._check(self.lib.sylph_update_code_generator(self.h))

    def update_code_generator_device(self, params: Dict[str, torch.Tensor]) -> None:
        """The same refresh straight from the optimiser's DEVICE tensors (sylph_update_code_generator_device): the operand
        layouts are packed by kernels on the current stream, nothing but the two scale scalars touches the host."""
        p_struct = self._codegen_tensor_struct(params)
        self._check(self.lib.sylph_update_code_generator_device(self.h, byref(p_struct), self._stream()))

    # ------------------------------------------------------------------ training backward (FCOS class tower)
    TOWER_PREFIX = "proposal_generator.fcos_head.cls_tower."

    def set_loss_box_branch(self, on: bool) -> None:
        """Off: fcos_loss_sums skips the box tower / predictors (the reference returns loss_fcos_cls alone when the box branch is frozen)."""
        self._check(self.lib.sylph_set_loss_box_branch(self.h, 1 if on else 0))

    def set_training(self, on: bool) -> None:
        """Keep the class tower's activations during the head pass of fcos_loss_sums (sylph_set_training)."""
        self._check(self.lib.sylph_set_training(self.h, 1 if on else 0))

    def _tower_tensor_struct(self, tensors: Dict[str, torch.Tensor]) -> TowerTensors:
        tt = TowerTensors()
        def ptr(name):
            t = tensors.get(self.TOWER_PREFIX + name)
            if t is None:
                return None
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), name
            return t.data_ptr()
        for i in range(min(int(self.mc.num_cls_convs), CG_MAX_TOWER)):
            tt.conv_w[i], tt.conv_b[i] = ptr(f"{3 * i}.weight"), ptr(f"{3 * i}.bias")
            tt.gn_w[i], tt.gn_b[i] = ptr(f"{3 * i + 1}.weight"), ptr(f"{3 * i + 1}.bias")
        return tt

    def cls_tower_backward(self, slot: int, codes: torch.Tensor, support_targets: Sequence[int], labels: torch.Tensor,
                           sums: torch.Tensor, params: Dict[str, torch.Tensor], global_pos_ctr: Optional[torch.Tensor] = None,
                           world_size: int = 1, grad_loss: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Gradients of loss_fcos_cls with respect to the class tower's tensors (sylph_cls_tower_backward); call after
        fcos_loss_sums(..., want_targets=True) with training mode on.  `codes`: the FINAL (C, 257) rows of that call."""
        codes = codes.to(self.device, torch.float32).contiguous()
        n_classes = codes.shape[0]
        st = (c_int64 * n_classes)(*[int(t) for t in support_targets])
        if grad_loss is not None:
            grad_loss = grad_loss.detach().to(self.device, torch.float32).reshape(1).contiguous()
        if not hasattr(self, "_lc"):
            self._lc = loss_config_from_cfg(self.cfg)
        params = {k: v for k, v in params.items() if k.startswith(self.TOWER_PREFIX)}
        grads = {k: torch.zeros_like(v) for k, v in params.items()}
        p_struct, g_struct = self._tower_tensor_struct(params), self._tower_tensor_struct(grads)
        self._check(self.lib.sylph_cls_tower_backward(
            self.h, slot, n_classes, c_void_p(codes.data_ptr()), st, byref(self._lc), c_void_p(labels.data_ptr()),
            c_void_p(sums.data_ptr()), c_void_p(global_pos_ctr.data_ptr()) if global_pos_ctr is not None else None, int(world_size),
            c_void_p(grad_loss.data_ptr()) if grad_loss is not None else None, byref(p_struct), byref(g_struct), self._stream()))
        self._keep_tbw = (codes, grad_loss, params)
        return grads

    def update_cls_tower_device(self, params: Dict[str, torch.Tensor]) -> None:
        """Re-prepare the class tower's weights from the optimiser's device tensors (sylph_update_cls_tower_device)."""
        p_struct = self._tower_tensor_struct(params)
        self._check(self.lib.sylph_update_cls_tower_device(self.h, byref(p_struct), self._stream()))

    def debug_read_buffer(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """Copy of one of the engine's named scratch buffers (debugging aid, sylph_debug_read_buffer)."""
        out = torch.empty(shape, device=self.device, dtype=dtype)
        self._check(self.lib.sylph_debug_read_buffer(self.h, name.encode(), c_void_p(out.data_ptr()), out.numel() * out.element_size(),
                                                     self._stream()))
        return out

    # ------------------------------------------------------------------ instrumentation
    def launch_count(self) -> int:
        return int(self.lib.sylph_launch_count(self.h))

    def set_profiling(self, on: bool) -> None:
        self._check(self.lib.sylph_set_profiling(self.h, int(on)))

    def timings(self, cap: int = 4096):
        names = ((c_char * 48) * cap)()
        ms = (c_float * cap)()
        fl = (c_double * cap)()
        by = (c_double * cap)()
        n = self.lib.sylph_get_timings(self.h, names, ms, fl, by, cap)
        return [(names[i].value.decode(), ms[i], fl[i], by[i]) for i in range(min(n, cap))]
