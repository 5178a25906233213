#!/usr/bin/env python3
"""Builds the overlay sources: copies of SEVEN reference files with the CUDA-backend hooks inserted, written to
integration/_build/src/ (git-ignored; nothing of the reference is committed to this repository).

Each hook is (file, anchor text that must occur exactly once [after an optional `after` marker], action).  The script
fails loudly when an anchor does not match, so a different reference version cannot be patched silently wrong.  It
also writes the result as a unified diff (integration/_build/overlay.patch) - the patch a maintainer would apply.

The hooks (all behind `#if SVT_CUDA`, all inert unless the environment sets SVT_CUDA=1):
  EbEncHandle.c                 svt_av1_enc_init: svt_cuda_backend_init() after setup_rtcd_internal (:1144-1145);
                                svt_av1_enc_deinit: svt_cuda_backend_deinit() (:1879)
  EbMotionEstimationProcess.c   the SB loop (:831-965) becomes the else-branch of svt_cuda_me_segment(...); the open-loop
                                intra search loop (:965-975) the else-branch of svt_cuda_ois_segment (SVT_CUDA_OIS=1, a
                                parity switch, off by default)
  EbCodingLoop.c                the per-SB deblocking of loop_filter_mode 1 (:3785-3795) is skipped when the frame is
                                deblocked on the GPU in dlf_kernel instead
  EbDlfProcess.c                svt_av1_pick_filter_level(FULL_IMAGE) + svt_av1_loop_filter_frame (:203-216) ->
                                svt_cuda_dlf_pick_frame / svt_cuda_dlf_frame; loop_filter_mode 1 pictures are
                                deblocked here, frame level, before the pre-CDEF preparation (:220)
  EbCdefProcess.c               cdef_seg_search of each segment (:510-515) skipped, and finish_cdef_search +
                                svt_av1_cdef_frame (:521-534) replaced by svt_cuda_cdef_picture for the whole picture
  EbRestProcess.c               svt_av1_loop_restoration_filter_frame (:533) -> svt_cuda_lr_frame (presets <= 6)
  EbPictureAnalysisProcess.c    compute_picture_spatial_statistics (:2929): the per-SB loop is skipped when
                                svt_cuda_pa_statistics did the picture (SVT_CUDA_PA=1, a parity switch, off by default)
"""
import difflib
import os
import sys

REF = os.environ.get("REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "src")

INCLUDE = '#if SVT_CUDA\n#include "svt_cuda_backend.h"\n#endif\n'

ME_CALL = (
    "#if SVT_CUDA\n"
    "                if (svt_cuda_me_segment(context_ptr, pcs_ptr, scs_ptr, pa_ref_obj_, input_padded_picture_ptr,\n"
    "                                        quarter_picture_ptr, sixteenth_picture_ptr, input_picture_ptr, segment_index,\n"
    "                                        x_sb_start_index, x_sb_end_index, y_sb_start_index, y_sb_end_index)) {\n"
    "                    /* the whole picture was searched on the GPU by the thread that got segment 0 */\n"
    "                } else\n"
    "#endif\n"
)

HOOKS = [
    # ---------------------------------------------------------------------------------------------------- EbEncHandle.c
    dict(file="Source/Lib/Encoder/Globals/EbEncHandle.c",
         anchor="    setup_rtcd_internal(enc_handle_ptr->scs_instance_array[0]->scs_ptr->static_config.use_cpu_flags);\n",
         after=None, action="insert_after",
         text="#if SVT_CUDA\n    svt_cuda_backend_init(); /* SVT_CUDA=1: picture-level stages go to libsvtav1_b200 */\n#endif\n"),
    dict(file="Source/Lib/Encoder/Globals/EbEncHandle.c",
         anchor="    if (handle) {\n", after="EB_API EbErrorType svt_av1_enc_deinit(EbComponentType *svt_enc_component){",
         action="insert_after", text="#if SVT_CUDA\n        svt_cuda_backend_deinit();\n#endif\n"),
    # ------------------------------------------------------------------------------------- EbMotionEstimationProcess.c
    dict(file="Source/Lib/Encoder/Codec/EbMotionEstimationProcess.c",
         anchor="                for (uint32_t y_sb_index = y_sb_start_index; y_sb_index < y_sb_end_index;\n",
         after="                use_scaled_source_refs_if_needed(pcs_ptr,\n", action="insert_before", text=ME_CALL),
    dict(file="Source/Lib/Encoder/Codec/EbMotionEstimationProcess.c",
         anchor="                for (uint32_t y_sb_index = y_sb_start_index; y_sb_index < y_sb_end_index;\n",
         after="                scs_ptr->static_config.enable_tpl_la)\n", action="insert_before",
         text="#if SVT_CUDA\n"
              "                if (svt_cuda_ois_segment(pcs_ptr, scs_ptr, input_picture_ptr, segment_index)) {\n"
              "                    /* every macroblock of the picture was searched on the GPU by the thread that got segment 0 */\n"
              "                } else\n"
              "#endif\n"),
    # --------------------------------------------------------------------------------------------------- EbCodingLoop.c
    dict(file="Source/Lib/Encoder/Codec/EbCodingLoop.c",
         anchor="    if (dlf_enable_flag && pcs_ptr->parent_pcs_ptr->loop_filter_mode == 1 && total_tile_cnt == 1) {\n",
         after="// First Pass Deblocking\n", action="replace",
         text="    if (dlf_enable_flag && pcs_ptr->parent_pcs_ptr->loop_filter_mode == 1 && total_tile_cnt == 1\n"
              "#if SVT_CUDA\n"
              "        && !svt_cuda_dlf_applies(pcs_ptr, scs_ptr) /* deblocked at frame level in dlf_kernel instead */\n"
              "#endif\n"
              "    ) {\n"),
    # --------------------------------------------------------------------------------------------------- EbDlfProcess.c
    dict(file="Source/Lib/Encoder/Codec/EbDlfProcess.c",
         anchor="            svt_av1_loop_filter_frame(recon_buffer, pcs_ptr, 0, 3);\n", after=None, action="replace",
         text="#if SVT_CUDA\n"
              "            if (svt_cuda_dlf_applies(pcs_ptr, scs_ptr))\n"
              "                svt_cuda_dlf_frame(pcs_ptr, scs_ptr, recon_buffer);\n"
              "            else\n"
              "#endif\n"
              "            svt_av1_loop_filter_frame(recon_buffer, pcs_ptr, 0, 3);\n"),
    dict(file="Source/Lib/Encoder/Codec/EbDlfProcess.c",
         anchor="            svt_av1_pick_filter_level(\n                context_ptr,\n"
                "                (EbPictureBufferDesc *)pcs_ptr->parent_pcs_ptr->enhanced_picture_ptr,\n"
                "                pcs_ptr,\n                LPF_PICK_FROM_FULL_IMAGE);\n",
         after=None, action="insert_before",
         text="#if SVT_CUDA\n"
              "            /* level search + deblocking as one GPU call (the svt_av1_loop_filter_frame hook below then returns) */\n"
              "            if (svt_cuda_dlf_applies(pcs_ptr, scs_ptr) && svt_cuda_dlf_pick_frame(pcs_ptr, scs_ptr, recon_buffer)) {\n"
              "            } else\n"
              "#endif\n"),
    dict(file="Source/Lib/Encoder/Codec/EbDlfProcess.c",
         anchor="        //pre-cdef prep\n", after=None, action="insert_before",
         text="#if SVT_CUDA\n"
              "        /* loop_filter_mode 1: av1_encode_decode skipped its per-SB deblocking; the frame is deblocked here */\n"
              "        if (dlf_enable_flag && pcs_ptr->parent_pcs_ptr->loop_filter_mode == 1 && total_tile_cnt == 1 &&\n"
              "            svt_cuda_dlf_applies(pcs_ptr, scs_ptr))\n"
              "            svt_cuda_dlf_frame(pcs_ptr, scs_ptr, NULL);\n"
              "#endif\n"),
    # -------------------------------------------------------------------------------------------------- EbCdefProcess.c
    dict(file="Source/Lib/Encoder/Codec/EbCdefProcess.c",
         anchor="        if (scs_ptr->seq_header.cdef_level && pcs_ptr->parent_pcs_ptr->cdef_level) {\n",
         after="void *cdef_kernel(void *input_ptr) {", action="insert_after",
         text="#if SVT_CUDA\n"
              "            if (svt_cuda_cdef_applies(pcs_ptr, scs_ptr)) {\n"
              "                /* searched for the whole picture by the thread that completes the last segment */\n"
              "            } else\n"
              "#endif\n"),
    dict(file="Source/Lib/Encoder/Codec/EbCdefProcess.c",
         anchor="                int32_t selected_strength_cnt[64] = {0};\n", after="void *cdef_kernel(void *input_ptr) {",
         action="insert_before",
         text="#if SVT_CUDA\n"
              "                if (svt_cuda_cdef_applies(pcs_ptr, scs_ptr))\n"
              "                    svt_cuda_cdef_picture(pcs_ptr, scs_ptr); /* search + finish_cdef_search + frame apply */\n"
              "                else\n"
              "#endif\n"
              "                {\n"),
    dict(file="Source/Lib/Encoder/Codec/EbCdefProcess.c",
         anchor="            } else {\n                frm_hdr->cdef_params.cdef_bits             = 0;\n",
         after="void *cdef_kernel(void *input_ptr) {", action="insert_before", text="                }\n"),
    # -------------------------------------------------------------------------------------------------- EbRestProcess.c
    dict(file="Source/Lib/Encoder/Codec/EbRestProcess.c",
         anchor="                    svt_av1_loop_restoration_filter_frame(cm->frame_to_show, cm, 0);\n",
         after="void *rest_kernel(void *input_ptr) {", action="insert_before",
         text="#if SVT_CUDA\n"
              "                    if (svt_cuda_lr_applies(pcs_ptr, scs_ptr))\n"
              "                        svt_cuda_lr_frame(pcs_ptr, scs_ptr); /* the frame apply; the search above stays the reference's */\n"
              "                    else\n"
              "#endif\n"),
    # --------------------------------------------------------------------------------------- EbPictureAnalysisProcess.c
    dict(file="Source/Lib/Encoder/Codec/EbPictureAnalysisProcess.c",
         anchor="    uint64_t pic_tot_variance = 0;\n", after="void compute_picture_spatial_statistics(SequenceControlSet", action="insert_after",
         text="#if SVT_CUDA\n"
              "    if (svt_cuda_pa_statistics(pcs_ptr, input_picture_ptr, input_padded_picture_ptr, sb_total_count)) return;\n"
              "#endif\n"),
]


def apply(src, hook):
    start = 0
    if hook["after"]:
        start = src.find(hook["after"])
        assert start >= 0, f"{hook['file']}: marker not found: {hook['after']!r}"
        assert src.find(hook["after"], start + 1) < 0, f"{hook['file']}: marker not unique: {hook['after']!r}"
    i = src.find(hook["anchor"], start)
    assert i >= 0, f"{hook['file']}: anchor not found: {hook['anchor']!r}"
    if not hook["after"]:
        assert src.find(hook["anchor"], i + 1) < 0, f"{hook['file']}: anchor not unique: {hook['anchor']!r}"
    j = i + len(hook["anchor"])
    if hook["action"] == "insert_before":
        return src[:i] + hook["text"] + src[i:]
    if hook["action"] == "insert_after":
        return src[:j] + hook["text"] + src[j:]
    if hook["action"] == "replace":
        return src[:i] + hook["text"] + src[j:]
    raise ValueError(hook["action"])


def main():
    os.makedirs(OUT, exist_ok=True)
    files = []
    for h in HOOKS:
        if h["file"] not in files:
            files.append(h["file"])
    patch = []
    for f in files:
        orig = open(os.path.join(REF, f), encoding="utf-8", errors="surrogateescape").read()
        new = orig
        for h in HOOKS:
            if h["file"] == f:
                new = apply(new, h)
        # the include goes after the first quoted include of the file
        k = new.find('#include "')
        k = new.find("\n", k) + 1
        new = new[:k] + INCLUDE + new[k:]
        dst = os.path.join(OUT, os.path.basename(f))
        old = open(dst, encoding="utf-8", errors="surrogateescape").read() if os.path.exists(dst) else None
        if old != new:  # keep timestamps stable for make
            open(dst, "w", encoding="utf-8", errors="surrogateescape").write(new)
        patch += difflib.unified_diff(orig.splitlines(True), new.splitlines(True), "a/" + f, "b/" + f, n=2)
    open(os.path.join(HERE, "_build", "overlay.patch"), "w", encoding="utf-8", errors="surrogateescape").writelines(patch)
    print("overlay: %d hooks in %d files -> %s" % (len(HOOKS), len(files), OUT))


if __name__ == "__main__":
    sys.exit(main())
