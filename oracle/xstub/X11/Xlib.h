/* Minimal stand-in for <X11/Xlib.h>: TEST INFRASTRUCTURE ONLY.
 *
 * The reference header (cpp_validation/taichi.h:17057-17059) includes X11
 * unconditionally on Linux and this image has no X11 headers.  The oracle
 * build never constructs a GUI, so the declarations below only have to let
 * taichi.h:17061-17142 compile; none of them is ever called. */
#ifndef MPM_ORACLE_XSTUB_XLIB_H
#define MPM_ORACLE_XSTUB_XLIB_H

typedef struct _XDisplayStub Display;
typedef struct _XVisualStub Visual;
typedef struct _XGCStub *GC;
typedef unsigned long Window;

struct XImage {
  char *data;
};

union XEvent {
  int type;
  long pad[24];
};

enum { ZPixmap = 2 };
enum { KeyPress = 2, ButtonPress = 4, Expose = 12 };
enum {
  KeyPressMask = 1L << 0,
  KeyReleaseMask = 1L << 1,
  ButtonPressMask = 1L << 2,
  ExposureMask = 1L << 15
};

static inline XImage *XCreateImage(Display *, Visual *, unsigned, int, int, char *data, unsigned, unsigned, int, int) {
  XImage *im = new XImage;
  im->data = data;
  return im;
}
static inline int XPending(Display *) { return 0; }
static inline int XNextEvent(Display *, XEvent *) { return 0; }
static inline Display *XOpenDisplay(const char *) { return nullptr; }
static inline Visual *DefaultVisual(void *, int) { return nullptr; }
static inline Window RootWindow(Display *, int) { return 0; }
static inline Window XCreateSimpleWindow(Display *, Window, int, int, unsigned, unsigned, unsigned, unsigned long,
                                         unsigned long) {
  return 0;
}
static inline int XSelectInput(Display *, Window, long) { return 0; }
static inline int XMapWindow(Display *, Window) { return 0; }
static inline GC DefaultGC(void *, int) { return nullptr; }
static inline int XPutImage(Display *, Window, GC, XImage *, int, int, int, int, unsigned, unsigned) { return 0; }
static inline int XStoreName(Display *, Window, const char *) { return 0; }

#endif
