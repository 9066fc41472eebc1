// TEST INFRASTRUCTURE: type stand-ins for <jack/jack.h> (no JACK server is involved offline).
#pragma once
#include <stdint.h>
typedef uint32_t jack_nframes_t;
typedef float jack_default_audio_sample_t;
typedef struct _jack_port jack_port_t;
typedef struct _jack_client jack_client_t;
