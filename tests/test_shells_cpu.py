"""The C/GLib element shells (gst-plugins-bad_b200/gst/*.c) compile: every plugin variant of gstb200vf.c and the
allocator / buffer pool, with `gcc -fsyntax-only -Wall -Werror` against the declaration-only headers of tests/stubs/
(no GLib or GStreamer in this image). Catches typos, wrong arities / types, missing struct members and
forgotten includes; the real build is gst/meson.build on a machine with GStreamer >= 1.19."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GST = os.path.join(ROOT, "gst-plugins-bad_b200", "gst")
FLAGS = ["gcc", "-std=gnu11", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-function",
         "-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + os.path.join(ROOT, "include"), "-I" + GST]
PLUGINS = ["bayer", "gaudieffects", "coloreffects", "geometrictransform", "videofiltersbad", "smooth", "videosignal"]


@pytest.mark.parametrize("plugin", PLUGINS)
def test_shell_compiles_for_every_plugin(plugin):
    r = subprocess.run(FLAGS + ["-DB200VF_PLUGIN=" + plugin, os.path.join(GST, "gstb200vf.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_allocator_and_pool_compile():
    r = subprocess.run(FLAGS + [os.path.join(GST, "gstb200vfmemory.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_the_check_has_teeth(tmp_path):
    """a call with a wrong arity, a missing struct member and an undeclared function are all rejected"""
    for body in ("gst_buffer_map (NULL, NULL);", "GstVideoFrame f; f.no_such_member = 0;", "gst_no_such_function (1);",
                 "b200vf_element_transform_host (NULL, NULL, NULL);"):
        src = tmp_path / "bad.c"
        src.write_text('#include <gst/gst.h>\n#include <gst/video/video.h>\n#include "b200vf.h"\nvoid f (void) { %s }\n' % body)
        r = subprocess.run(FLAGS + [str(src)], capture_output=True, text=True)
        assert r.returncode != 0, body


def test_meson_builds_every_plugin_of_the_factory_table(vf):
    """gst/meson.build's plugin list == the plugins named by the library's introspection table"""
    meson = open(os.path.join(GST, "meson.build")).read()
    plugins = {f["plugin"] for f in vf.factories().values()}
    for p in plugins:
        assert "'%s'" % p in meson, p
    assert plugins == set(PLUGINS)
