"""Host-side restatement of the reference's chunk loop as scan units (grab.cc:151-159) for tests:
a file larger than chunk_size becomes windows [off, off+clen), off += chunk_size - 4096."""


def windows(size, chunk_size=1 << 30, overlap=0x1000):
    off = 0
    out = []
    while off < size:
        clen = min(chunk_size, size - off)
        out.append((off, clen))
        off += chunk_size - overlap
    return out
