// fake_x11.h -- TEST INFRASTRUCTURE.  The handful of Xlib types (public ABI layouts of <X11/Xlib.h> / Xlibint.h / XShm) that Mesa's xlib GLX
// front end touches when it creates an off-screen context; used by fake_x11.c (the stand-in libX11.so.6) and gl_harness.c.  Nothing here talks to
// an X server: there is none.
#pragma once
#include <stddef.h>
typedef unsigned long XID, VisualID, Window, Drawable, Colormap, Pixmap, Font;
typedef char* XPointer;
typedef int Bool, Status;
typedef struct _XExtData XExtData;
typedef struct _XGC* GC;
typedef struct { XExtData* ext_data; VisualID visualid; int c_class; unsigned long red_mask, green_mask, blue_mask; int bits_per_rgb; int map_entries; } Visual;
typedef struct { int depth; int nvisuals; Visual* visuals; } Depth;
struct _XDisplay;
typedef struct {
  XExtData* ext_data; struct _XDisplay* display; Window root; int width, height; int mwidth, mheight; int ndepths; Depth* depths; int root_depth;
  Visual* root_visual; GC default_gc; Colormap cmap; unsigned long white_pixel; unsigned long black_pixel; int max_maps, min_maps; int backing_store;
  Bool save_unders; long root_input_mask;
} Screen;
typedef struct { XExtData* ext_data; int depth; int bits_per_pixel; int scanline_pad; } ScreenFormat;
typedef struct _XDisplay {
  XExtData* ext_data; void* private1; int fd; int private2; int proto_major_version; int proto_minor_version; char* vendor;
  XID private3, private4, private5; int private6; XID (*resource_alloc)(struct _XDisplay*);
  int byte_order; int bitmap_unit; int bitmap_pad; int bitmap_bit_order; int nformats; ScreenFormat* pixmap_format; int private8; int release;
  void *private9, *private10; int qlen; unsigned long last_request_read; unsigned long request; XPointer private11, private12, private13, private14;
  unsigned max_request_size; void* db; int (*private15)(struct _XDisplay*); char* display_name; int default_screen; int nscreens; Screen* screens;
  unsigned long motion_buffer; unsigned long private16; int min_keycode; int max_keycode; XPointer private17, private18; int private19; char* xdefaults;
  char* scratch_buffer; unsigned long scratch_length; int ext_number; struct _XExten* ext_procs;   /* Xlibint.h */
  char pad[8192];
} Display;
typedef struct { Visual* visual; VisualID visualid; int screen; int depth; int c_class; unsigned long red_mask, green_mask, blue_mask; int colormap_size; int bits_per_rgb; } XVisualInfo;
typedef struct _XImage {
  int width, height; int xoffset; int format; char* data; int byte_order; int bitmap_unit; int bitmap_bit_order; int bitmap_pad; int depth;
  int bytes_per_line; int bits_per_pixel; unsigned long red_mask, green_mask, blue_mask; XPointer obdata;
  struct { struct _XImage* (*create_image)(); int (*destroy_image)(struct _XImage*); unsigned long (*get_pixel)(struct _XImage*, int, int);
           int (*put_pixel)(struct _XImage*, int, int, unsigned long); struct _XImage* (*sub_image)(); int (*add_pixel)(struct _XImage*, long); } f;
} XImage;

typedef struct { int extension, major_opcode, first_event, first_error; } XExtCodes;
typedef struct _XExten { struct _XExten* next; XExtCodes codes; void *create_GC, *copy_GC, *flush_GC, *free_GC, *create_Font, *free_Font, *close_display, *error, *error_string;
  char* name; void* error_values; void* before_flush; struct _XExten* next_flush; } _XExtension;
