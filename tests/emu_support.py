"""Test-side switch to the CPU emulation of the kernel source (tools/cpu_emu).  The product has
one library and no notion of an emulation (solver/b200.py: load_library() takes no path and the
device-pointer API insists on CUDA tensors); the tests that run the kernel source on the CPU
swap the loaded library object and relax those two checks HERE, and restore them afterwards."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, 'tools', 'cpu_emu')
EMU_LIB = os.path.join(EMU_DIR, '_build', 'libomgb200_emu.so')


def build():
    src = [os.path.join(ROOT, 'omg_tools_b200', 'csrc', f) for f in ('omg_b200.cu', 'omg_sp.cuh', 'omg_sp_host.cuh')]
    src += [os.path.join(ROOT, 'include', 'omg_b200.h'), os.path.join(EMU_DIR, 'cuda_runtime.h'),
            os.path.join(EMU_DIR, 'emu_runtime.cpp')]
    if (not os.path.exists(EMU_LIB) or
            any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in src)):
        subprocess.check_call([os.path.join(EMU_DIR, 'build.sh')])


def _check_any_tensors(tensors, lib=None):
    import torch
    for t in tensors:
        if t.dtype != torch.float64 or not t.is_contiguous():
            raise ValueError('expected contiguous float64 tensors')
    return tensors[0].is_cuda


def _check_any_int_tensors(tensors):
    import torch
    for t in tensors:
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise ValueError('expected contiguous int32 tensors')


def _stream_or_none(on_gpu, device, stream):
    return None


def activate():
    """Point omg_tools_b200.solver.b200 at the emulation library; returns the state to restore."""
    from omg_tools_b200.solver import b200
    build()
    saved = (b200._lib, b200._check_device_tensors, b200._check_int_tensors, b200._stream_handle)
    b200._lib = b200.bind(C.CDLL(EMU_LIB))
    b200._check_device_tensors = _check_any_tensors
    b200._check_int_tensors = _check_any_int_tensors
    b200._stream_handle = _stream_or_none
    return saved


def restore(saved):
    from omg_tools_b200.solver import b200
    b200._lib, b200._check_device_tensors, b200._check_int_tensors, b200._stream_handle = saved
