"""A real ECMAScript engine for the reference's JavaScript -- TEST INFRASTRUCTURE, never imported by fspt_b200/.

The image has no Node / browser, but Nsight Compute's host directory ships Qt 6 with its QML JavaScript engine
(QJSEngine, "V4": ES2016 + modules, IEEE doubles, typed arrays).  It has no headers and no Python binding, so it is
driven through ctypes on the Itanium-ABI symbols of libQt6Core / libQt6Qml: QString / QJSValue are returned through a
hidden first pointer, `this` comes next.  libQt6Core links glib for an event dispatcher that is never used here (no
event loop); the image has no glib, so two stand-in libraries with the 16 imported names are compiled on first use
into oracle/_ref/qt_stubs/ and preloaded.

    eng = JSEngine()                      # raises JSUnavailable when Qt's libraries are not there
    eng.import_module("/path/bvh.js", "B")   # ES module with its relative imports -> global `B`
    text = eng.evaluate("JSON.stringify(new B.BoundingBox().min)")
"""
import ctypes as C
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "_ref", "qt_stubs")
_GLIB_NAMES = ("g_main_context_default g_main_context_iteration g_main_context_new g_main_context_pop_thread_default "
               "g_main_context_push_thread_default g_main_context_ref g_main_context_unref g_main_context_wakeup "
               "g_source_add_poll g_source_attach g_source_destroy g_source_new g_source_remove_poll "
               "g_source_set_can_recurse g_source_set_name g_source_unref").split()


class JSUnavailable(RuntimeError):
    pass


def qt_dir():
    cands = [os.environ.get("FSPT_QT_DIR", "")] + sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/linux-desktop-*-x64"))
    for d in cands:
        if d and os.path.exists(os.path.join(d, "libQt6Qml.so.6")) and os.path.exists(os.path.join(d, "libQt6Core.so.6")):
            return d
    return None


def _have_lib(name):
    try:
        C.CDLL(name)
        return True
    except OSError:
        return False


def _stub(name, body):
    os.makedirs(_STUBS, exist_ok=True)
    so = os.path.join(_STUBS, name)
    if not os.path.exists(so):
        src = so + ".c"
        with open(src, "w") as f:
            f.write(body)
        # the SONAME is what lets the loader accept the preloaded stand-in for libQt6Core's DT_NEEDED entry
        subprocess.check_call(["gcc", "-shared", "-fPIC", "-Wl,-soname," + name, "-o", so, src])
    return so


class _QString(C.Structure):  # QArrayDataPointer<char16_t>
    _fields_ = [("d", C.c_void_p), ("ptr", C.c_void_p), ("size", C.c_longlong)]


_state = {}


def _load():
    if _state:
        return _state
    d = qt_dir()
    if d is None:
        raise JSUnavailable("no Qt 6 with QJSEngine found (looked under /opt/nvidia/nsight-compute/*/host and $FSPT_QT_DIR)")
    try:
        if not _have_lib("libglib-2.0.so.0"):
            body = "#include <stdlib.h>\n" + "".join("void %s(void) { abort(); }\n" % n for n in _GLIB_NAMES)
            C.CDLL(_stub("libglib-2.0.so.0", body), mode=C.RTLD_GLOBAL)
            C.CDLL(_stub("libgthread-2.0.so.0", "void fspt_gthread_stand_in(void) {}\n"), mode=C.RTLD_GLOBAL)
        os.environ.setdefault("QT_NO_GLIB", "1")
        core = C.CDLL(os.path.join(d, "libQt6Core.so.6"), mode=C.RTLD_GLOBAL)
        C.CDLL(os.path.join(d, "libQt6Network.so.6"), mode=C.RTLD_GLOBAL)
        qml = C.CDLL(os.path.join(d, "libQt6Qml.so.6"), mode=C.RTLD_GLOBAL)
    except (OSError, subprocess.CalledProcessError) as e:
        raise JSUnavailable("Qt's JavaScript engine could not be loaded: %s" % e)
    argc = C.c_int(1)
    argv = (C.c_char_p * 2)(b"fspt-js", None)
    app = C.create_string_buffer(64)
    core._ZN16QCoreApplicationC1ERiPPci(app, C.byref(argc), argv, C.c_int(0x060000))
    _state.update(core=core, qml=qml, keep=(argc, argv, app))
    return _state


def available():
    try:
        _load()
        return True
    except JSUnavailable:
        return False


class JSError(RuntimeError):
    pass


class JSEngine:
    def __init__(self):
        st = _load()
        self.core, self.qml = st["core"], st["qml"]
        self.eng = C.create_string_buffer(64)
        self.qml._ZN9QJSEngineC1Ev(self.eng)
        # console.log etc. (QJSEngine::ConsoleExtension = 0x2), into a fresh QJSValue() = the global object
        undefined = C.create_string_buffer(16)
        self.qml._ZN8QJSValueC1ENS_12SpecialValueE(undefined, C.c_int(1))
        self.qml._ZN9QJSEngine17installExtensionsE6QFlagsINS_9ExtensionEERK8QJSValue(self.eng, C.c_int(0x2), undefined)

    def _qstr(self, s):
        b = s.encode("utf-8")
        q = _QString()
        f = self.core._ZN7QString8fromUtf8E14QByteArrayView
        f.restype = None
        f(C.byref(q), C.c_longlong(len(b)), C.c_char_p(b))
        return q

    def _to_py(self, val):
        out = _QString()
        g = self.qml._ZNK8QJSValue8toStringEv
        g.restype = None
        g(C.byref(out), val)
        text = C.string_at(out.ptr, out.size * 2).decode("utf-16-le") if out.ptr else ""
        iserr = self.qml._ZNK8QJSValue7isErrorEv
        iserr.restype = C.c_bool
        if iserr(val):
            raise JSError(text)
        return text

    def evaluate(self, src, name="<fspt>"):
        """Runs a script; returns String(result).  A thrown exception becomes JSError."""
        val = C.create_string_buffer(16)
        p, n = self._qstr(src), self._qstr(name)
        f = self.qml._ZN9QJSEngine8evaluateERK7QStringS2_iP5QListIS0_E
        f.restype = None
        f(val, self.eng, C.byref(p), C.byref(n), C.c_int(1), None)
        return self._to_py(val)

    def import_module(self, path, global_name):
        """import * as <global_name> from '<path>' (an ES module on disk; its relative imports are followed)."""
        ns = C.create_string_buffer(16)
        p = self._qstr(path)
        f = self.qml._ZN9QJSEngine12importModuleERK7QString
        f.restype = None
        f(ns, self.eng, C.byref(p))
        self._to_py(ns)  # raises when the module failed to load
        glob_ = C.create_string_buffer(16)
        g = self.qml._ZNK9QJSEngine12globalObjectEv
        g.restype = None
        g(glob_, self.eng)
        name = self._qstr(global_name)
        s = self.qml._ZN8QJSValue11setPropertyERK7QStringRKS_
        s.restype = None
        s(glob_, C.byref(name), ns)
