"""The oracle is test infrastructure: nothing under arpeggio_b200/ may import, load or execute it,
and the product has no CPU fallback path."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'arpeggio_b200')


def _sources():
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.c', '.cpp')) or f == 'Makefile':
                yield os.path.join(base, f)


def test_product_never_touches_the_oracle():
    pat = re.compile(r'(from\s+oracle|import\s+oracle|liborc|arp_oracle|orc_pairs|orc_classify)')
    for path in _sources():
        text = open(path).read()
        assert not pat.search(text), f'{path} refers to the oracle'


def test_product_has_no_torch_or_triton():
    pat = re.compile(r'^\s*(import|from)\s+(torch|triton)\b', re.M)
    for path in _sources():
        if path.endswith('.py'):
            assert not pat.search(open(path).read()), f'{path} imports torch/triton'


def test_header_lives_outside_oracle():
    assert os.path.exists(os.path.join(ROOT, 'include', 'arpeggio_cuda.h'))
    assert 'NOT product code' in open(os.path.join(ROOT, 'oracle', 'arp_oracle.c')).read()
