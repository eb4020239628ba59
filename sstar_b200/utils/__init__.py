from .utils import (  # noqa: F401
    align_population_data_by_position,
    check_anc_allele,
    filter_data,
    parse_ind_file,
    read_anc_allele,
    read_data,
    read_geno_data,
    split_genome,
    validate_unique_variant_positions,
)
