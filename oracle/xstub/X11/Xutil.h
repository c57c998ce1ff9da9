/* Empty stand-in for <X11/Xutil.h> (see Xlib.h in this directory). TEST INFRASTRUCTURE ONLY. */
