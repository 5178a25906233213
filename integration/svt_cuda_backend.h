/* svt_cuda_backend.h - entry points the overlay hooks call (see svt_cuda_backend.c, overlay.py). */
#ifndef SVT_CUDA_BACKEND_H
#define SVT_CUDA_BACKEND_H
#include <stdint.h>

#include "EbPictureControlSet.h"
#include "EbSequenceControlSet.h"
#include "EbReferenceObject.h"
#include "EbMotionEstimationProcess.h"

void svt_cuda_backend_init(void);   /* svt_av1_enc_init, after setup_rtcd_internal (EbEncHandle.c:1144-1145) */
void svt_cuda_backend_deinit(void); /* svt_av1_enc_deinit (EbEncHandle.c:1879) */

/* stage predicates: must give the same answer at every hook of one picture */
int svt_cuda_me_active(void);
int svt_cuda_dlf_applies(PictureControlSet *pcs_ptr, SequenceControlSet *scs_ptr);
int svt_cuda_cdef_applies(PictureControlSet *pcs_ptr, SequenceControlSet *scs_ptr);
int svt_cuda_lr_applies(PictureControlSet *pcs_ptr, SequenceControlSet *scs_ptr);

int svt_cuda_me_segment(MotionEstimationContext_t *context_ptr, PictureParentControlSet *pcs_ptr, SequenceControlSet *scs_ptr,
                        EbPaReferenceObject *pa_ref_obj, EbPictureBufferDesc *input_padded_picture_ptr,
                        EbPictureBufferDesc *quarter_picture_ptr, EbPictureBufferDesc *sixteenth_picture_ptr,
                        EbPictureBufferDesc *input_picture_ptr, uint32_t segment_index, uint32_t x_sb_start_index,
                        uint32_t x_sb_end_index, uint32_t y_sb_start_index, uint32_t y_sb_end_index);
void svt_cuda_dlf_frame(PictureControlSet *pcs_ptr, SequenceControlSet *scs_ptr, EbPictureBufferDesc *recon_buffer);
int  svt_cuda_dlf_pick_frame(PictureControlSet *pcs_ptr, SequenceControlSet *scs_ptr, EbPictureBufferDesc *recon_buffer);
void svt_cuda_cdef_picture(PictureControlSet *pcs_ptr, SequenceControlSet *scs_ptr);
void svt_cuda_lr_frame(PictureControlSet *pcs_ptr, SequenceControlSet *scs_ptr);

int  svt_cuda_pa_statistics(PictureParentControlSet *pcs_ptr, EbPictureBufferDesc *input_picture_ptr,
                            EbPictureBufferDesc *input_padded_picture_ptr, uint32_t sb_total_count); /* SVT_CUDA_PA=1 */

int  svt_cuda_ois_segment(PictureParentControlSet *pcs_ptr, SequenceControlSet *scs_ptr, EbPictureBufferDesc *input_picture_ptr,
                          uint32_t segment_index); /* SVT_CUDA_OIS=1 */

/* SVT_CUDA_PROFILE=1: wall time the stage threads spend in each stage, CPU path included (stage: 0 me, 1 dlf, 2 cdef) */
int64_t svt_cuda_prof_begin(void);
void    svt_cuda_prof_end_cpu(int stage, int64_t t0);
#endif
