#!/usr/bin/env python3
"""Mirror the run-time DATA assets the hot path needs out of a Mitsuba source/install tree into refdata/.

The reference resolves these through its FileResolver at run time (data/microfacet/*.dat, src/bsdfs/rtrans.h:95-97)
or compiles them into plugin binaries (Hosek-Wilkie RGB coefficients, src/emitters/sunsky/skymodeldata.h; CIE 1931
colour-matching tables, src/libcore/spectrum.cpp:743-1141).  They are data, not code; refdata/ is git-ignored and is
re-created by `__graft_entry__.build()` whenever the reference tree is present (it travels to the GPU box with the
snapshot, like built .so files).

Layout written:
  refdata/microfacet/{beckmann,ggx,phong}.dat      byte copies
  refdata/sunsky/hosek_rgb.f64                     3 channels x (1080 config + 120 radiance) float64, little endian
  refdata/cie1931.f32                              4 x 471 float32: wavelengths, X, Y, Z
  refdata/sobol.bin                                the direction-number tables compiled into the `sobol` sampler plugin (src/samplers/sobolseq.cpp: Joe & Kuo's numbers
                                                   as 1024 x 52 32-bit matrices, and the van-der-Corput / Sobol enumeration matrices with their inverses for m = 1..26):
                                                   uint32 magic 'SOBL', dims, size, rowsVdc, rowsInv; matrices32[dims*size] uint32; vdc[rowsVdc*size] uint64; vdc_inv[rowsInv*size] uint64
"""
import os, re, shutil, struct, sys

def parse_c_array(text, name):
    m = re.search(r'\b' + re.escape(name) + r'\s*\[[^\]]*\]\s*=\s*\{(.*?)\};', text, re.S)
    if not m:
        raise RuntimeError('array %s not found' % name)
    body = re.sub(r'//.*?$|/\*.*?\*/', '', m.group(1), flags=re.S | re.M)
    return [float(tok.rstrip('fF')) for tok in re.findall(r'[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?[fF]?', body)]

def main(ref, out):
    os.makedirs(os.path.join(out, 'microfacet'), exist_ok=True)
    os.makedirs(os.path.join(out, 'sunsky'), exist_ok=True)
    for n in ('beckmann', 'ggx', 'phong'):
        shutil.copyfile(os.path.join(ref, 'data', 'microfacet', n + '.dat'), os.path.join(out, 'microfacet', n + '.dat'))
    sky = open(os.path.join(ref, 'src', 'emitters', 'sunsky', 'skymodeldata.h')).read()
    with open(os.path.join(out, 'sunsky', 'hosek_rgb.f64'), 'wb') as f:
        for c in (1, 2, 3):
            cfg = parse_c_array(sky, 'datasetRGB%d' % c)
            rad = parse_c_array(sky, 'datasetRGBRad%d' % c)
            assert len(cfg) == 1080 and len(rad) == 120, (len(cfg), len(rad))
            f.write(struct.pack('<%dd' % len(cfg), *cfg))
            f.write(struct.pack('<%dd' % len(rad), *rad))
    spec = open(os.path.join(ref, 'src', 'libcore', 'spectrum.cpp')).read()
    with open(os.path.join(out, 'cie1931.f32'), 'wb') as f:
        for name in ('CIE_wavelengths', 'CIE_X_entries', 'CIE_Y_entries', 'CIE_Z_entries'):
            v = parse_c_array(spec, name)
            assert len(v) == 471, (name, len(v))
            f.write(struct.pack('<471f', *v))
    sob = open(os.path.join(ref, 'src', 'samplers', 'sobolseq.cpp')).read()
    def block(name):
        i = sob.index('Matrices::' + name); i = sob.index('{', sob.index('=', i))
        j = sob.index('\n};', i)
        return sob[i + 1:j + 2]
    m32 = [int(t.rstrip('U'), 16) for t in re.findall(r'0x[0-9a-fA-F]+U?\b', block('matrices32['))]
    assert len(m32) == 1024 * 52, len(m32)
    def rows(name):
        out_rows = []
        for body in re.findall(r'\{\s*// m = \d+(.*?)\}', block(name), re.S):
            v = [int(t[:-3], 16) for t in re.findall(r'0x[0-9a-fA-F]+ULL', body)]
            assert 0 < len(v) <= 52
            out_rows.append(v + [0] * (52 - len(v)))
        return out_rows
    vdc, inv = rows('vdc_sobol_matrices[]'), rows('vdc_sobol_matrices_inv[]')
    assert len(vdc) >= 16 and len(inv) >= 16, (len(vdc), len(inv))
    with open(os.path.join(out, 'sobol.bin'), 'wb') as f:
        f.write(struct.pack('<5I', 0x4c424f53, 1024, 52, len(vdc), len(inv)))
        f.write(struct.pack('<%dI' % len(m32), *m32))
        for table in (vdc, inv):
            for r in table:
                f.write(struct.pack('<52Q', *r))
    print('refdata written to', out)

if __name__ == '__main__':
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'refdata')
    main(ref, out)
