/* TEST INFRASTRUCTURE: stands in for FreeType's <ft2build.h> (font_gl.h only needs the types to exist). */
#ifndef FAKE_FT2BUILD_H
#define FAKE_FT2BUILD_H
#define FT_FREETYPE_H "ft_fake.h"
#endif
