"""Corrupt the *inflated* BAM stream (record fields: block_size, l_read_name, n_cigar_op, l_seq, random spans, truncation),
re-wrap it as BGZF and run `inquistr-b200 bamstat` on it: the host reader must fail cleanly (exit 1 with a message) or
succeed, never crash. Meant for a sanitizer build of the CLI (use a SMALL BAM, the whole stream is rewritten per case):
  g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=all -pthread -o /tmp/cli_asan \
      inquistr_b200/csrc/host/*.cpp -Linquistr_b200/lib -linqcall -lz -Wl,-rpath,$PWD/inquistr_b200/lib
  python tools/fuzz_bam.py small.bam /tmp/cli_asan 400"""
import gzip, struct, subprocess, sys, zlib, random, os
src, exe, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = gzip.open(src, 'rb').read()   # BGZF is multi-member gzip
def bgzf(data, level=1):
    out = bytearray()
    for i in range(0, len(data), 0xff00):
        chunk = data[i:i + 0xff00]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = len(comp) + 25
        out += struct.pack('<BBBBIBBHBBHH', 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, bsize)
        out += comp + struct.pack('<II', zlib.crc32(chunk) & 0xffffffff, len(chunk))
    out += bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000')
    return bytes(out)
# header length
l_text = struct.unpack_from('<i', raw, 4)[0]
p = 8 + l_text
n_ref = struct.unpack_from('<i', raw, p)[0]; p += 4
for _ in range(n_ref):
    l = struct.unpack_from('<i', raw, p)[0]; p += 4 + l + 4
first = p
rng = random.Random(7)
crashes = 0
codes = {}
for it in range(n):
    b = bytearray(raw)
    mode = rng.randrange(6)
    # walk to a random record
    q = first; k = rng.randrange(200)
    for _ in range(k):
        if q + 4 > len(b): break
        bs = struct.unpack_from('<i', b, q)[0]
        if q + 4 + bs > len(b): break
        q += 4 + bs
    if q + 40 > len(b): q = first
    if mode == 0: struct.pack_into('<i', b, q, rng.choice([-1, 0, 3, 31, 2**31 - 1, rng.randrange(1, 1 << 20)]))          # block_size
    elif mode == 1: b[q + 12] = rng.randrange(256)                                                                      # l_read_name
    elif mode == 2: struct.pack_into('<H', b, q + 16, rng.choice([0, 1, 65535, rng.randrange(65536)]))                   # n_cigar_op
    elif mode == 3: struct.pack_into('<i', b, q + 20, rng.choice([-5, 0, 2**31 - 1, rng.randrange(1 << 24)]))           # l_seq
    elif mode == 4:
        a = rng.randrange(first, len(b) - 64)
        for j in range(a, a + rng.randrange(1, 64)): b[j] = rng.randrange(256)
    else: b = b[:rng.randrange(first, len(b))]                                                                          # truncated stream
    path = os.path.join(os.path.dirname(os.path.abspath(src)), "fz.bam")
    open(path, 'wb').write(bgzf(bytes(b)))
    r = subprocess.run([exe, 'bamstat', path], capture_output=True, timeout=120)
    codes[r.returncode] = codes.get(r.returncode, 0) + 1
    if r.returncode not in (0, 1, 101) or b'AddressSanitizer' in r.stderr or b'runtime error' in r.stderr:
        crashes += 1
        print('CRASH mode', mode, 'rc', r.returncode, r.stderr[-600:].decode(errors='replace'))
        if crashes > 3: break
print('done', n, 'cases; exit codes', codes, 'crashes', crashes)
