"""Host mirror of ORBVocabulary (orb_slam2/include/ORBVocabulary.h:31 = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>)
over the C ABI: loadFromTextFile + transform.  The tree descent (k Hamming distances per level) and the assembly of the
BowVector / FeatureVector run on the GPU (csrc/orb_bow.cu); there is no CPU path."""
import ctypes as C
import os

import numpy as np

from . import _lib

# DBoW2 enums (BowVector.h:36-53)
TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = 0, 1, 2, 3, 4, 5


class ORBVocabulary:
    def __init__(self, handle, device):
        self._h = handle
        self.device = device
        v = [C.c_int32() for _ in range(6)]
        _lib.check(_lib.lib().orb_voc_info(self._h, *[C.byref(x) for x in v]))
        self.k, self.L, self.n_nodes, self.n_words, self.scoring, self.weighting = [int(x.value) for x in v]

    @classmethod
    def from_arrays(cls, k, L, scoring, weighting, parent, is_leaf, desc, weight, device=0):
        parent = np.ascontiguousarray(parent, np.int32)
        is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8)
        weight = np.ascontiguousarray(weight, np.float64)
        h = C.c_void_p()
        _lib.check(_lib.lib().orb_voc_create(C.byref(h), device, k, L, scoring, weighting, len(parent), _lib.ptr(parent),
                                             _lib.ptr(is_leaf), _lib.ptr(desc), _lib.ptr(weight)))
        return cls(h, device)

    @classmethod
    def loadFromTextFile(cls, path, device=0):
        h = C.c_void_p()
        _lib.check(_lib.lib().orb_voc_load_text(C.byref(h), device, os.fsencode(path)))
        return cls(h, device)

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.lib().orb_voc_destroy(self._h)
            self._h = None

    def empty(self):
        return self.n_nodes <= 1

    def size(self):
        return self.n_words

    def transform_features(self, desc, levelsup=0):
        """transform(feature, id, weight, &nid, levelsup) per descriptor -> (word_id, weight, node_id)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word = np.zeros(n, np.int32); weight = np.zeros(n, np.float64); node = np.zeros(n, np.int32)
        _lib.check(_lib.lib().orb_bow_transform_features(self._h, _lib.ptr(desc), n, levelsup, _lib.ptr(word), _lib.ptr(weight),
                                                         _lib.ptr(node)))
        return word, weight, node

    def transform_batch(self, descs, levelsup=4):
        """transform(features, BowVector, FeatureVector, levelsup) for a list of descriptor arrays.  Returns one
        ((bow_word, bow_value), (fv_node, fv_start, fv_feat)) pair per frame, fv in CSR form."""
        descs = [np.ascontiguousarray(d, np.uint8).reshape(-1, 32) for d in descs]
        nf = len(descs)
        off = np.zeros(nf + 1, np.int32)
        off[1:] = np.cumsum([len(d) for d in descs])
        n = int(off[-1])
        allq = np.concatenate(descs) if n else np.zeros((0, 32), np.uint8)
        bn = np.zeros(nf, np.int32); bw = np.zeros(n, np.int32); bv = np.zeros(n, np.float64)
        fn = np.zeros(nf, np.int32); fnode = np.zeros(n, np.int32); fs = np.zeros(n + nf, np.int32); ff = np.zeros(n, np.int32)
        _lib.check(_lib.lib().orb_bow_transform(self._h, _lib.ptr(allq), _lib.ptr(off), nf, levelsup, _lib.ptr(bn), _lib.ptr(bw),
                                                _lib.ptr(bv), _lib.ptr(fn), _lib.ptr(fnode), _lib.ptr(fs), _lib.ptr(ff)))
        out = []
        for f in range(nf):
            o = int(off[f]); nb = int(bn[f]); nn = int(fn[f])
            st = fs[o + f:o + f + nn + 1].copy()
            out.append(((bw[o:o + nb].copy(), bv[o:o + nb].copy()), (fnode[o:o + nn].copy(), st, ff[o:o + int(st[nn])].copy())))
        return out

    def transform(self, desc, levelsup=4):
        return self.transform_batch([desc], levelsup)[0]
