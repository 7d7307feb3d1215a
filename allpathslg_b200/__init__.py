"""allpathslg_b200 -- B200-native k-mer spectrum engine for ALLPATHS-LG's k-mer hot path."""
from .kmers import (ApgkError, KmerCounter, KmerFreqTable, KmerGroup, KmerParcelsBuilder, KmerSpectrum, SortKmers, owner_of,
                    synth_params, words_per_kmer)

__all__ = ["ApgkError", "KmerCounter", "KmerFreqTable", "KmerGroup", "KmerParcelsBuilder", "KmerSpectrum", "SortKmers", "owner_of",
           "synth_params", "words_per_kmer"]
