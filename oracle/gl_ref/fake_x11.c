// Minimal stand-in for libX11 / libXext: just enough for Mesa's xlib GLX state tracker to create an off-screen (pbuffer) context.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fake_x11.h"
#define LOG(...) do { if (getenv("FAKEX_DEBUG")) { fprintf(stderr, "[fakex] " __VA_ARGS__); fputc('\n', stderr); } } while (0)
static Visual g_visual = {0, 0x21, 4 /*TrueColor*/, 0xFF0000, 0x00FF00, 0x0000FF, 8, 256};
void* _Xglobal_lock = 0;
void (*_XLockMutex_fn)(void*) = 0;
void (*_XUnlockMutex_fn)(void*) = 0;
Visual* fakex_visual(void) { return &g_visual; }
XVisualInfo* XGetVisualInfo(Display* d, long mask, XVisualInfo* t, int* n) {
  LOG("XGetVisualInfo mask=%lx depth=%d class=%d id=%lx", mask, t->depth, t->c_class, t->visualid);
  *n = 0;
  if ((mask & 0x4) && t->depth != 24) return 0;
  if ((mask & 0x8) && t->c_class != 4) return 0;
  if ((mask & 0x1) && t->visualid != g_visual.visualid) return 0;
  XVisualInfo* v = calloc(1, sizeof *v);
  v->visual = &g_visual; v->visualid = g_visual.visualid; v->screen = 0; v->depth = 24; v->c_class = 4;
  v->red_mask = g_visual.red_mask; v->green_mask = g_visual.green_mask; v->blue_mask = g_visual.blue_mask; v->colormap_size = 256; v->bits_per_rgb = 8;
  *n = 1;
  return v;
}
int XFree(void* p) { free(p); return 1; }
static int destroy_image(XImage* im) { free(im); return 1; }
XImage* XCreateImage(Display* d, Visual* v, unsigned depth, int format, int offset, char* data, unsigned w, unsigned h, int pad, int bpl) {
  LOG("XCreateImage depth=%u %ux%u", depth, w, h);
  XImage* im = calloc(1, sizeof *im);
  im->width = w; im->height = h; im->format = format; im->data = data; im->byte_order = 0; im->bitmap_unit = 32; im->bitmap_pad = pad; im->depth = depth;
  im->bits_per_pixel = depth > 16 ? 32 : depth > 8 ? 16 : 8; im->bytes_per_line = bpl ? bpl : (int)w * im->bits_per_pixel / 8;
  im->red_mask = 0xFF0000; im->green_mask = 0xFF00; im->blue_mask = 0xFF; im->f.destroy_image = destroy_image;
  return im;
}
XImage* XShmCreateImage() { LOG("XShmCreateImage"); return 0; }
int XShmAttach() { return 0; }
int XShmPutImage() { return 0; }
int XQueryExtension(Display* d, const char* name, int* a, int* b, int* c) { LOG("XQueryExtension %s", name); return 0; }
XExtCodes* XAddExtension(Display* d) {
  LOG("XAddExtension");
  _XExtension* e = calloc(1, sizeof *e);
  e->codes.extension = d->ext_number++;
  e->next = d->ext_procs;
  d->ext_procs = e;
  return &e->codes;
}
Colormap XCreateColormap() { return 0x31; }
GC XCreateGC() { LOG("XCreateGC"); return (GC)calloc(1, 128); }
int XFreeGC(Display* d, GC gc) { free(gc); return 1; }
Pixmap XCreatePixmap() { return 0x41; }
int XFreePixmap() { return 1; }
int XDrawString16() { return 1; }
int XFillRectangle() { return 1; }
int XFlush() { return 1; }
int XFreeFontInfo() { return 1; }
Status XGetGeometry(Display* d, Drawable dr, Window* root, int* x, int* y, unsigned* w, unsigned* h, unsigned* bw, unsigned* depth) {
  LOG("XGetGeometry"); if (root) *root = 1; if (x) *x = 0; if (y) *y = 0; if (w) *w = 64; if (h) *h = 64; if (bw) *bw = 0; if (depth) *depth = 24; return 1; }
XImage* XGetImage() { return 0; }
Status XGetWindowAttributes() { LOG("XGetWindowAttributes"); return 0; }
int XPutImage() { return 1; }
void* XQueryFont() { return 0; }
void* XSetErrorHandler(void* h) { return 0; }
int XSetForeground() { return 1; }
int XSetFunction() { return 1; }
int XSync() { return 1; }
void* XSynchronize() { return 0; }
