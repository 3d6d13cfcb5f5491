"""Genome container (host plumbing).  Same surface as the reference's catch/genome.py:9-143 for
what the hot path reads: `.seqs` (list of str), `.chrs`, `.size()`."""
from collections import OrderedDict


class Genome:
    def __init__(self, seqs, chrs=None):
        if len(seqs) > 1 and chrs is None:
            raise ValueError("When there is more than one sequence, chrs should also be specified")
        self.seqs = seqs
        self.chrs = chrs
        self._size = None
        self._hash = None

    def divided_into_chrs(self):
        return len(self.seqs) > 1

    def size(self, only_unambig=False):
        if only_unambig:
            return sum(s.count(b) for s in self.seqs for b in 'ATCG')
        if self._size is None:
            self._size = sum(len(s) for s in self.seqs)
        return self._size

    def break_into_fragments(self, fragment_length, include_full_end=False):
        def pieces(seq):
            for i in range(0, len(seq), fragment_length):
                frag = seq[i:i + fragment_length]
                if include_full_end and len(frag) < fragment_length:
                    frag = seq[max(0, len(seq) - fragment_length):]
                yield frag
        out = OrderedDict()
        if self.chrs is None:
            for i, frag in enumerate(pieces(self.seqs[0])):
                out[str(i)] = frag
        else:
            for name, seq in self.chrs.items():
                for i, frag in enumerate(pieces(seq)):
                    out[name + '-' + str(i)] = frag
        return Genome.from_chrs(out)

    def __hash__(self):
        if self._hash is None:
            self._hash = hash(tuple(self.seqs))
        return self._hash

    def __eq__(self, other):
        return isinstance(other, Genome) and self.seqs == other.seqs and self.chrs == other.chrs

    @staticmethod
    def from_chrs(seqs_by_chr):
        for s in seqs_by_chr.values():
            if not isinstance(s, str):
                raise TypeError("Sequences must be strings")
        return Genome(list(seqs_by_chr.values()), seqs_by_chr)

    @staticmethod
    def from_one_seq(seq):
        if not isinstance(seq, str):
            raise TypeError("seq must be a string")
        return Genome([seq])
