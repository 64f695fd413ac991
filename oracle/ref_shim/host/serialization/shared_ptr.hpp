// Stand-in: serialization of shared pointers is not used in oracle/_ref.
