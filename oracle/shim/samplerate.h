#pragma once
typedef struct SRC_STATE_tag SRC_STATE;
typedef struct { const float* data_in; float* data_out; long input_frames, output_frames, input_frames_used, output_frames_gen; int end_of_input; double src_ratio; } SRC_DATA;
enum { SRC_SINC_BEST_QUALITY = 0, SRC_SINC_MEDIUM_QUALITY = 1, SRC_SINC_FASTEST = 2 };
