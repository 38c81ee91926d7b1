"""Output file-name templates of the all-pairs path (reference ``src/utility/constant.py:4-8``)."""
FUNUNIFRAC_OUT__MAIN_FILE_NAME = 'fununifrac_out_{0}.main.npy'
FUNUNIFRAC_OUT__MAIN_BASIS__FILE_NAME = 'fununifrac_out_{0}.main.basis.npy'
FUNUNIFRAC_OUT__DIFFAB_FILE_NAME = 'fununifrac_out_{0}.diffab'
FUNUNIFRAC_OUT__DIFFAB_NODES_FILE_NAME = 'fununifrac_out_{0}.diffab.nodes.npy'
FUNUNIFRAC_OUT__PUSH_FILE_NAME = 'fununifrac_out_{0}.push.pkl'
