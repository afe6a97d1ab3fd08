int mo_dummy;
