#pragma once
typedef struct SNDFILE_tag SNDFILE;
typedef struct { long frames; int samplerate, channels, format, sections, seekable; } SF_INFO;
