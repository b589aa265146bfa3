/* TEST INFRASTRUCTURE: see ft2build.h in this directory. */
#ifndef FAKE_FT_H
#define FAKE_FT_H
typedef struct FT_LibraryRec_ *FT_Library;
typedef struct FT_FaceRec_ *FT_Face;
typedef struct FT_GlyphSlotRec_ *FT_GlyphSlot;
#endif
